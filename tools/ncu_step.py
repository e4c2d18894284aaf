"""Eager (host-launched) training steps of one BASELINE configuration for ncu: python tools/ncu_step.py [config] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import multivae_b200 as mb  # noqa: E402
from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "ns"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = bench.CONFIGS[config]["batch"]
dev = torch.device("cuda", 0)
model = bench.build_model(config, dev)
model.compute_dtype = torch.bfloat16
host = bench.synthetic_batch(B, config=config)
tr = BaseTrainer(model, mb.MultimodalBaseDataset(data=host), training_config=BaseTrainerConfig(per_device_train_batch_size=B, learning_rate=1e-3))
model.train()
res = mb.DatasetOutput(data={k: v.to(dev) for k, v in host.items()})
kw = dict(bench.config_spec(config).get("fwd", {}))
for _ in range(steps):
    tr.step_batch(res, **kw)
torch.cuda.synchronize()
print("done")
