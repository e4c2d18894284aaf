"""The oracle PORT (oracle/port) against golden vectors produced by the REAL reference
(oracle/make_golden.py): loss, loss_sum, metrics, per-term lw tensors and every parameter gradient."""
import glob
import os

import pytest
import torch

from oracle.cases import CASES
from oracle.replay import run_port

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order only


def _close(a, b, rtol=RTOL, atol=1e-6):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (a.flatten()[:4], b.flatten()[:4])


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference_golden(name):
    rec = torch.load(os.path.join(GOLD, f"elbo_{name}.pt"), weights_only=False)
    details = {}
    loss, loss_sum, metrics, p = run_port(CASES[name], rec, details=details)
    _close(loss.detach(), rec["loss"])
    _close(loss_sum.detach(), rec["loss_sum"])
    for k, v in rec["metrics"].items():
        _close(torch.as_tensor(metrics[k]).detach(), v, rtol=1e-4)
    if "lws" in rec:
        for m, lw in rec["lws"].items():
            _close(details[m]["lw"].detach(), lw, rtol=1e-5, atol=1e-4)
    scale = max(1.0, float(rec["loss"].abs()))
    for k, g in rec["grads"].items():
        pg = p[k].grad
        if g is None:
            assert pg is None or float(pg.abs().sum()) == 0.0, k
            continue
        assert pg is not None, k
        assert abs(float(pg.double().sum()) - g["sum"]) <= 1e-4 * max(g["abssum"], 1e-3) + 1e-5, k
        _close(pg.flatten()[:8], g["head"], rtol=1e-3, atol=1e-5 * scale)
        if g["full"] is not None:
            _close(pg, g["full"], rtol=1e-3, atol=1e-5 * scale)


def test_port_fp64_agrees():
    """fp64 replay of the port stays within 1e-5 relative of the fp32 reference value."""
    name = "mmvaeplus_dreg"
    rec = torch.load(os.path.join(GOLD, f"elbo_{name}.pt"), weights_only=False)
    loss, *_ = run_port(CASES[name], rec, dtype=torch.float64, want_grads=False)
    assert abs(float(loss) - float(rec["loss"])) <= 1e-5 * abs(float(rec["loss"]))


def test_cfg5_fp32_conditioning():
    """BASELINE config 5 (3x64x64 image, K = 10, DReG) is the worst-conditioned case: an fp64 evaluation of the same algorithm
    differs from the reference's fp32 value by ~7e-5 relative (the loss is a nearly one-hot weighted sum of log-weights of
    magnitude 1e4).  The GPU path is still held to 1e-4 against the reference's fp32 value (precise expf / logf in the
    latent kernel, tests/gpu_checks.py)."""
    name = "cfg5_mmvaeplus_celeba"
    rec = torch.load(os.path.join(GOLD, f"elbo_{name}.pt"), weights_only=False)
    loss, *_ = run_port(CASES[name], rec, dtype=torch.float64, want_grads=False)
    r = abs(float(loss) - float(rec["loss"])) / abs(float(rec["loss"]))
    assert 2e-5 < r < 5e-4, r


def test_golden_files_complete():
    names = {os.path.basename(p)[5:-3] for p in glob.glob(os.path.join(GOLD, "elbo_*.pt"))}
    assert names == set(CASES)
