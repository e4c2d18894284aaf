"""Network restatements (oracle/port/nets.py) vs outputs of the REAL reference modules
(tests/golden/nets_*.pt), plus the reference's own known-answer facts for this path."""
import os

import pytest
import torch

from oracle.port import elbo as E
from oracle.port import nets as N

GOLD = os.path.join(os.path.dirname(__file__), "golden")

FWD = {
    "enc_resnet_mmnist": lambda p, x: dict(zip(("embedding", "log_covariance", "style_embedding", "style_log_covariance"), N.encoder_resnet_mmnist(p, "", x))),
    "dec_resnet_mmnist": lambda p, z: {"reconstruction": N.decoder_resnet_mmnist(p, "", z)},
    "enc_conv_mmnist": lambda p, x: dict(zip(("embedding", "log_covariance"), N.encoder_conv_mmnist_adapted(p, "", x))),
    "dec_conv_mmnist": lambda p, z: {"reconstruction": N.decoder_conv_mmnist(p, "", z)},
    "enc_svhn": lambda p, x: dict(zip(("embedding", "log_covariance"), N.encoder_vae_svhn(p, "", x))),
    "dec_svhn": lambda p, z: {"reconstruction": N.decoder_vae_svhn(p, "", z)},
    "enc_mlp": lambda p, x: dict(zip(("embedding", "log_covariance"), N.encoder_vae_mlp(p, "", x))),
    "enc_mlp_style": lambda p, x: dict(zip(("embedding", "log_covariance", "style_embedding", "style_log_covariance"), N.encoder_vae_mlp_style(p, "", x))),
    "dec_mlp": lambda p, z: {"reconstruction": N.decoder_ae_mlp(p, "", z, (1, 28, 28))},
    "enc_cub_resnet": lambda p, x: dict(zip(("embedding", "log_covariance"), N.cub_resnet_encoder(p, "", x))),
    "dec_cub_resnet": lambda p, z: {"reconstruction": N.cub_resnet_decoder(p, "", z)},
}


def golden_input(rec):
    g = torch.Generator().manual_seed(77)
    x = torch.rand(rec["in_shape"], generator=g) if rec["kind"] == "x" else torch.randn(rec["in_shape"], generator=g)
    return x, g


@pytest.mark.parametrize("name", sorted(FWD))
def test_net_port_matches_reference(name):
    rec = torch.load(os.path.join(GOLD, f"nets_{name}.pt"), weights_only=False)
    p = {k: v.clone().requires_grad_(True) for k, v in N.synth_state_dict(rec["state_shapes"], seed=rec["sd_seed"]).items()}
    x, _ = golden_input(rec)
    x.requires_grad_(True)
    out = FWD[name](p, x)
    for k, v in rec["outputs"].items():
        assert torch.allclose(out[k], v, rtol=1e-4, atol=1e-5), (name, k)
    sum((out[k] * rec["cot"][k]).sum() for k in out).backward()
    assert torch.allclose(x.grad, rec["grad_in"], rtol=1e-3, atol=1e-5)
    for k, g in rec["grads"].items():
        assert torch.allclose(p[k].grad.flatten()[:8], g["head"], rtol=1e-3, atol=1e-4), (name, k)
        assert abs(float(p[k].grad.double().sum()) - g["sum"]) <= 1e-4 * max(g["abssum"], 1e-3) + 1e-5


def test_mopoe_subset_table_known_answers():
    # reference tests/test_mopoe.py:136-145: 3 modalities -> 8 keys including "", in this order
    s = E.mopoe_subsets(["mod1", "mod2", "mod3"])
    assert list(s.keys()) == ["", "mod1", "mod2", "mod3", "mod1_mod2", "mod1_mod3", "mod2_mod3", "mod1_mod2_mod3"]
    assert E.mopoe_subset_bitmasks(["a", "b", "c"]) == [1, 2, 4, 3, 5, 6, 7]
    assert len(E.mopoe_subset_bitmasks([f"m{i}" for i in range(5)])) == 31


def test_mopoe_deterministic_selection_known_answers():
    # B=256,S=31 -> [8]*30+[16]; B=32 -> [1]*30+[2]; exact multiples split evenly
    def sizes(B, S):
        idx = E.mopoe_sample_to_subset(B, S)
        assert bool((idx[1:] >= idx[:-1]).all())
        return torch.bincount(idx.long(), minlength=S).tolist()

    assert sizes(256, 31) == [8] * 30 + [16]
    assert sizes(32, 31) == [1] * 30 + [2]
    for B in (31, 62, 93, 310):
        assert sizes(B, 31) == [B // 31] * 31
    assert sizes(6, 7) == [0] * 6 + [6]


def test_poe_matches_closed_form():
    mus = torch.tensor([[[1.0]], [[3.0]]])
    lvs = torch.zeros(2, 1, 1)
    mu, lv = E.poe(mus, lvs, eps=0.0)
    assert torch.allclose(mu, torch.tensor([[2.0]])) and torch.allclose(lv, torch.log(torch.tensor([[0.5]])))
    mu2, lv2 = E.stable_poe(mus, lvs)
    assert torch.allclose(mu2, mu) and torch.allclose(lv2, lv)
    one = E.stable_poe(mus[:1], lvs[:1])
    assert one[0] is not None and torch.equal(one[0], mus[0])
