"""Parity of the CUDA path (through the C-ABI) with the reference's golden vectors and the oracle port:
loss / loss_sum / metrics within 1e-4 relative, per-term lw tensors, every parameter gradient, exact-zero
gradients for fully masked modalities, MoPoE subset indices bit-exact."""
import pytest
import torch

from oracle.cases import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_case_matches_reference(name):
    from tests.gpu_checks import check_case
    check_case(name, verbose=True)


def test_mopoe_selection_bit_exact():
    from oracle.port.elbo import mopoe_sample_to_subset, mopoe_subset_bitmasks
    from tests.gpu_checks import run_product
    out, model, rec = run_product("mopoe_5mod")
    sel = model._last["sel"].cpu()
    assert torch.equal(sel, mopoe_sample_to_subset(40, 31))
    assert model._last["subsets"].cpu().tolist() == mopoe_subset_bitmasks([f"m{i}" for i in range(5)])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dist,D", [("laplace", 2352), ("normal", 3072), ("bernoulli", 784), ("laplace", 10), ("normal", 12288)])
def test_lpx_kernels_vs_torch(dist, D, dtype):
    """mv_moe_lpx_fwd / bwd against torch.distributions on the same (rounded) inputs, incl. ragged D."""
    import torch.distributions as td
    from multivae_b200 import _cabi as C
    g = torch.Generator(device="cuda").manual_seed(1)
    Cn, K, B = 3, 4, 5
    recon = torch.rand(Cn, K, B, D, device="cuda", generator=g).to(dtype)
    x = torch.rand(B, D, device="cuda", generator=g)
    if dist == "bernoulli":
        recon = ((recon.float() - 0.5) * 6).to(dtype)
        x = (x > 0.5).float()
    mask = torch.tensor([1, 0, 1, 1, 1], dtype=torch.uint8, device="cuda")
    lpx = torch.empty(Cn, K, B, device="cuda")
    lib = C.lib()
    C.check(lib.mv_moe_lpx_fwd(C.ptr(recon), C.dtype_code(recon), C.ptr(x), C.ptr(lpx), Cn, K, B, D, C.DIST[dist], 0.75, 1.7, C.ptr(mask), 0, C.stream()), "fwd")
    r32 = recon.float().requires_grad_(True)
    d = {"laplace": lambda: td.Laplace(r32, 0.75), "normal": lambda: td.Normal(r32, 0.75), "bernoulli": lambda: td.Bernoulli(logits=r32)}[dist]()
    ref = (d.log_prob(x).mul(1.7).sum(-1)) * mask.float()
    assert torch.allclose(lpx, ref, rtol=2e-5, atol=1e-3), float((lpx - ref).abs().max())
    coef = torch.randn(Cn, K, B, device="cuda", generator=g)
    gl = torch.tensor([0.5], device="cuda")
    gr = torch.empty_like(recon)
    C.check(lib.mv_moe_lpx_bwd(C.ptr(recon), C.dtype_code(recon), C.ptr(x), C.ptr(coef), C.ptr(gl), C.ptr(gr), Cn, K, B, D, C.DIST[dist], 0.75, 1.7, C.ptr(mask), C.stream()), "bwd")
    (ref * coef).sum().mul(0.5).backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert torch.allclose(gr.float(), r32.grad, rtol=tol, atol=tol * float(r32.grad.abs().max()))


def test_missing_library_fails_loudly(monkeypatch):
    from multivae_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libmultivae_b200.so")
    with pytest.raises(_cabi.NativeLibraryError):
        _cabi.lib()
