"""2-GPU check (torchrun): the early all-reduce of the decoder gradients (overlapped with the encoders' backward) gives the same
parameters as the single all-reduce after the backward pass, eager and CUDA-graph mode.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_overlap_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import multivae_b200 as mb  # noqa: E402
from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
B = 8
host = {k: v[: B] + 0.01 * rank for k, v in bench.synthetic_batch(B).items()}
results = {}
for graph in (False, True):
    for overlap in (False, True):
        torch.manual_seed(7)
        model = bench.build_model("ns", dev)
        model.compute_dtype = torch.bfloat16
        model.model_config.K = 2
        tr = BaseTrainer(model, mb.MultimodalBaseDataset(data=host), training_config=BaseTrainerConfig(
            per_device_train_batch_size=B, learning_rate=0.0, world_size=world, rank=rank, local_rank=lr, use_cuda_graph=graph,
            graph_warmup_steps=1, overlap_allreduce=overlap))
        model.train()
        torch.manual_seed(100 + rank)
        batch = mb.DatasetOutput(data={k: v.to(dev) for k, v in host.items()})
        for _ in range(4):
            out = tr.step_batch(batch)
        torch.cuda.synchronize()
        # lr = 0: the parameters stay put, so the all-reduced gradient of the last step is comparable across the four modes
        # (under Adam a sign flip of a near-zero gradient element moves a parameter by +-lr: parameters are not comparable)
        results[(graph, overlap)] = (float(out.loss_sum.detach()), tr.flat.flat.detach().clone())
        fired = tr._comm_stream is not None
        assert fired == overlap, (graph, overlap, fired)
ref = results[(False, False)]
for key, (loss, flat) in results.items():
    err = float((flat - ref[1]).norm() / ref[1].norm())
    # replicas identical across ranks
    other = flat.clone()
    dist.broadcast(other, src=0)
    assert float((other - flat).abs().max()) == 0.0, ("replicas diverged", key)
    print(f"rank {rank} graph={key[0]} overlap={key[1]} loss {loss:.4f} relative L2 of the reduced gradient vs plain {err:.3e}", flush=True)
    assert err <= 1e-3, (key, err)
if rank == 0:
    print("overlap check ok", flush=True)
del tr, model, out, results   # CUDA graphs that captured NCCL kernels must go before the communicator does
import gc  # noqa: E402
gc.collect()
torch.cuda.synchronize()
import threading  # noqa: E402
threading.Timer(30.0, lambda: os._exit(0)).start()
dist.destroy_process_group()
os._exit(0)
