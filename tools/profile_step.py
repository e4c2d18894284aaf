"""torch.profiler view of one training step (which non-native kernels are left)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import multivae_b200 as mb
from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
model = bench.north_star_model(dev)
model.compute_dtype = torch.bfloat16
host = bench.synthetic_batch(B, pinned=True)
tr = BaseTrainer(model, mb.MultimodalBaseDataset(data=host), training_config=BaseTrainerConfig(per_device_train_batch_size=B, learning_rate=1e-3, optimizer_cls="Adam"))
res = mb.DatasetOutput(data={k: v.to(dev) for k, v in host.items()})
for _ in range(3):
    tr.step_batch(res)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step_batch(res)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
