"""One eager (host-launched) north-star training step for ncu: python tools/ncu_step.py [B] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import multivae_b200 as mb
from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
model = bench.north_star_model(dev)
model.compute_dtype = torch.bfloat16
host = bench.synthetic_batch(B)
tr = BaseTrainer(model, mb.MultimodalBaseDataset(data=host), training_config=BaseTrainerConfig(per_device_train_batch_size=B, learning_rate=1e-3))
res = mb.DatasetOutput(data={k: v.to(dev) for k, v in host.items()})
for _ in range(steps):
    tr.step_batch(res)
torch.cuda.synchronize()
print("done")
