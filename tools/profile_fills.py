"""Which Python call sites launch the small ATen kernels of one eager training step (torch.profiler with stacks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import multivae_b200 as mb
from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
model = bench.north_star_model(dev)
model.compute_dtype = torch.bfloat16
host = bench.synthetic_batch(B, pinned=True)
tr = BaseTrainer(model, mb.MultimodalBaseDataset(data=host), training_config=BaseTrainerConfig(per_device_train_batch_size=B, learning_rate=1e-3, optimizer_cls="Adam"))
res = mb.DatasetOutput(data={k: v.to(dev) for k, v in host.items()})
for _ in range(3):
    tr.step_batch(res, allow_graph=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr.step_batch(res, allow_graph=False)
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_stack_n=6)
rows = [e for e in ka if e.key in ("aten::zeros", "aten::fill_", "aten::zero_", "aten::zeros_like", "aten::new_zeros", "aten::add_", "aten::add", "aten::neg", "aten::mul", "aten::copy_", "aten::to", "aten::_to_copy")]
rows.sort(key=lambda e: -e.count)
for e in rows[:40]:
    stack = [s for s in e.stack if "multivae_b200" in s or "torch/autograd" in s or "optim" in s][:3]
    print(f"{e.key:18s} x{e.count:5d}  {' <- '.join(s.split('/')[-1] for s in stack)}")
