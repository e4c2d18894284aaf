"""Host logic of the data-parallel training step on CPU (gloo, world_size 2): batch sharding equals
DistributedSampler, the flat-gradient all-reduce reproduces DDP's gradient averaging, and `train_step` returns the
reference's (epoch_loss, metrics) conventions.  The stand-in model is plain torch (the real models need the CUDA
library and are covered by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

import multivae_b200 as mb
from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig, FlatGrads, shard_indices


class TinyModel(nn.Module):
    """forward(inputs) -> ModelOutput(loss = SUM over the batch, like MMVAE/MMVAE+)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.lin = nn.Linear(6, 4)
        self.device = None

    def forward(self, inputs, **kwargs):
        x = inputs.data["a"]
        loss = (self.lin(x) ** 2).sum()
        return mb.ModelOutput(loss=loss, loss_sum=loss, metrics={"m": loss.detach() * 0 + 2.0})

    def update(self):
        pass


def _dataset(n=16):
    g = torch.Generator().manual_seed(11)
    return mb.MultimodalBaseDataset(data={"a": torch.randn(n, 6, generator=g)})


def test_shard_indices_match_distributed_sampler():
    from torch.utils.data import DistributedSampler
    for n, w in [(16, 2), (17, 4), (5, 8), (256, 8)]:
        ds = list(range(n))
        for r in range(w):
            ref = list(DistributedSampler(ds, num_replicas=w, rank=r, shuffle=False))
            assert shard_indices(n, w, r) == ref
            # the reference's trainer keeps the sampler's default shuffle=True: same permutation per (seed, epoch)
            for epoch in (0, 3):
                smp = DistributedSampler(ds, num_replicas=w, rank=r, shuffle=True, seed=5)
                smp.set_epoch(epoch)
                assert shard_indices(n, w, r, shuffle=True, seed=5, epoch=epoch) == list(smp)


def test_flat_grads_are_views_and_survive_backward():
    m = TinyModel()
    f = FlatGrads(m.parameters())
    out = m(_dataset(4)[slice(0, 4)] if False else mb.DatasetOutput(data={"a": torch.ones(4, 6)}))
    out.loss.backward()
    f.rebind()
    assert f.flat.abs().sum() > 0
    for p in m.parameters():
        assert p.grad.data_ptr() >= f.flat.data_ptr()
    f.zero()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in m.parameters())


def test_train_step_single_process_conventions():
    ds = _dataset(16)
    tr = BaseTrainer(TinyModel(), ds, training_config=BaseTrainerConfig(per_device_train_batch_size=4, no_cuda=True, learning_rate=1e-2))
    w0 = tr.model.lin.weight.detach().clone()
    loss, metrics = tr.train_step(epoch=1)
    assert isinstance(loss, float) and loss > 0 and metrics == {"m": 2.0}
    assert not torch.equal(w0, tr.model.lin.weight)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    ds = _dataset(16)
    cfg = BaseTrainerConfig(per_device_train_batch_size=8, no_cuda=True, learning_rate=0.05, optimizer_cls="SGD")
    tr = BaseTrainer(TinyModel(), ds, training_config=cfg)
    batch = next(iter(tr.local_batches()))
    out = tr.step_batch(batch)
    # plain lists: tensors in a Queue travel as shared-memory handles that die with the worker
    q.put((rank, tr.flat.flat.tolist(), tr.model.lin.weight.detach().reshape(-1).tolist(), batch.data["a"].tolist(),
           float(out.loss.detach())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gradient_average_matches_manual():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (_, g0, w0, x0, _), (_, g1, w1, x1, _) = [tuple(torch.tensor(v) if isinstance(v, list) else v for v in r) for r in res]
    assert torch.equal(g0, g1) and torch.equal(w0, w1)          # replicas stay identical
    ds = _dataset(16)
    from torch.utils.data import DistributedSampler
    for r, x in ((0, x0), (1, x1)):   # DistributedSampler shards (shuffle=True, seed 0, epoch 0), first batch of 8
        idx = list(DistributedSampler(list(range(16)), num_replicas=2, rank=r, shuffle=True, seed=0))[:8]
        assert torch.equal(x, ds.data["a"][torch.tensor(idx)])
    # manual DDP semantics: mean over ranks of each rank's own loss gradient
    m = TinyModel()
    gs = []
    for x in (x0, x1):
        m.zero_grad()
        (m.lin(x) ** 2).sum().backward()
        gs.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]))
    assert torch.allclose(g0, (gs[0] + gs[1]) / 2, rtol=1e-6, atol=1e-6)
