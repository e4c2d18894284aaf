"""use_cuda_graph=True for every model, past the warm-up: the captured step must replay (no host->device copies or host
randomness inside the step) and follow the same parameter trajectory as the eager step on the same inputs and noise."""
import copy

import pytest
import torch

import multivae_b200 as mb
from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig

pytestmark = pytest.mark.gpu

DIMS = {"a": (3, 8, 8), "b": (1, 6, 6), "c": (10,)}
CFGS = {
    "mvtcae": (mb.MVTCAE, lambda: mb.MVTCAEConfig(n_modalities=3, latent_dim=6, input_dims=dict(DIMS))),
    "mvae": (mb.MVAE, lambda: mb.MVAEConfig(n_modalities=3, latent_dim=6, input_dims=dict(DIMS), k=0, warmup=4)),
    "mopoe": (mb.MoPoE, lambda: mb.MoPoEConfig(n_modalities=3, latent_dim=6, input_dims=dict(DIMS), beta=2.5)),
    "mmvae": (mb.MMVAE, lambda: mb.MMVAEConfig(n_modalities=3, latent_dim=6, input_dims=dict(DIMS), K=3)),
    "mmvaeplus": (mb.MMVAEPlus, lambda: mb.MMVAEPlusConfig(n_modalities=3, latent_dim=6, modalities_specific_dim=4, input_dims=dict(DIMS), K=3)),
}


class FixedNoise:
    """The same standard draws for the i-th request of every step (persistent device tensors: valid under graph replay)."""

    def __init__(self):
        self.bank, self.i = {}, 0

    def begin(self):
        self.i = 0

    def __call__(self, shape, kind, dev):
        key = (self.i, tuple(shape))
        self.i += 1
        if key not in self.bank:
            g = torch.Generator().manual_seed(100 + key[0])
            self.bank[key] = mb.elbo.standard_noise(tuple(shape), kind, "cpu", generator=g).to(dev)
        return self.bank[key]


def _run(name, graph, steps=7):
    cls, mk = CFGS[name]
    torch.manual_seed(0)
    model = cls(mk())
    noise = FixedNoise()
    model.noise_source = noise
    B = 16
    data = {m: torch.rand(B, *d, generator=torch.Generator().manual_seed(i)) for i, (m, d) in enumerate(DIMS.items())}
    ds = mb.MultimodalBaseDataset(data=data)
    cfg = BaseTrainerConfig(per_device_train_batch_size=B, learning_rate=1e-3, use_cuda_graph=graph, graph_warmup_steps=2, shuffle=False)
    tr = BaseTrainer(model, ds, training_config=cfg)
    batch = mb.DatasetOutput(data={m: t.cuda() for m, t in data.items()})
    losses = []
    for i in range(steps):
        noise.begin()
        out = tr.step_batch(batch, epoch=1 + i // 3, batch_ratio=(i % 3) / 3)   # MVAE: the KL weight changes every step
        losses.append(float(out.loss_sum.detach()))
    return losses, copy.deepcopy({k: v.detach().cpu() for k, v in model.state_dict().items()}), tr


@pytest.mark.parametrize("name", sorted(CFGS))
def test_graph_step_matches_eager_step(name):
    le, se, _ = _run(name, graph=False)
    lg, sg, tr = _run(name, graph=True)
    assert any("graph" in st for st in tr._graphs.values()), "the step was never captured"
    assert len(tr._graphs) == 1, "one signature, one graph"
    assert all(l == l for l in lg)
    for a, b in zip(le, lg):
        assert abs(a - b) <= 1e-4 * abs(a) + 1e-4, (name, le, lg)
    for k in se:
        assert torch.allclose(se[k], sg[k], rtol=1e-3, atol=1e-5), (name, k)


def test_train_step_with_prefetched_batches_matches_plain_steps():
    """train_step double-buffers the host->device copies (prefetch of batch i+1 during batch i): same trajectory as feeding the
    batches one by one, eager and graph mode."""
    cls, mk = CFGS["mvtcae"]
    data = {m: torch.rand(48, *d, generator=torch.Generator().manual_seed(i)).pin_memory() for i, (m, d) in enumerate(DIMS.items())}
    finals = []
    for graph, via_train_step in ((False, False), (False, True), (True, True)):
        torch.manual_seed(0)
        model = cls(mk())
        noise = FixedNoise()
        model.noise_source = lambda shape, kind, dev, n=noise: (n.begin(), n(shape, kind, dev))[1]   # the same draw every step
        ds = mb.MultimodalBaseDataset(data=data)
        tr = BaseTrainer(model, ds, training_config=BaseTrainerConfig(per_device_train_batch_size=16, learning_rate=1e-3, use_cuda_graph=graph,
                                                                      graph_warmup_steps=1, shuffle=False))
        for epoch in (1, 2):
            if via_train_step:
                loss, _ = tr.train_step(epoch)
                assert loss == loss
            else:
                for i in range(3):
                    tr.step_batch(mb.DatasetOutput(data={m: t[16 * i:16 * i + 16] for m, t in data.items()}), epoch=epoch)
        finals.append({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    for other in finals[1:]:
        for k in finals[0]:
            assert torch.allclose(finals[0][k], other[k], rtol=1e-3, atol=1e-5), k
