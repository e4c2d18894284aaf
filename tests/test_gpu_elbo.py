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


def test_north_star_golden_on_the_bf16_tensor_core_path():
    """MMVAE+ / ResNet golden of the REAL reference (K = 10, B = 4, non-initial weights) through the native tcgen05 encoders and
    decoders: loss and every-parameter gradient error against the fp32 reference, bounded at 3x the measured values."""
    from tests.gpu_checks import NS_BF16_GRAD_L2_TOL, NS_BF16_LOSS_TOL, check_case_bf16
    e = check_case_bf16("ns_mmvaeplus_resnet", verbose=True)
    assert e["loss"] <= NS_BF16_LOSS_TOL, e
    assert e["grad_rel_l2"] <= NS_BF16_GRAD_L2_TOL, e


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


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dist,D", [("laplace", 2352), ("normal", 784), ("laplace", 12288)])
def test_lpx_batched_launch_equals_per_modality_launches(dist, D, dtype):
    """mv_moe_lpx_fwd_multi / mv_moe_lpx_bwd_multi (all reconstructed modalities in one launch) against the
    per-modality kernels on the same inputs: lpx within fp32 summation-order noise, g_recon bit-identical."""
    from multivae_b200 import _cabi as C
    g = torch.Generator(device="cuda").manual_seed(7)
    Cn, K, B, R = 3, 4, 6, 3
    lib = C.lib()
    recons = [torch.rand(Cn, K, B, D, device="cuda", generator=g).to(dtype) for _ in range(R)]
    xs = [torch.rand(B, D, device="cuda", generator=g) for _ in range(R)]
    masks = [torch.tensor([1, 1, 0, 1, 1, 1], dtype=torch.uint8, device="cuda"), None, torch.ones(B, dtype=torch.uint8, device="cuda")]
    scales, rescales = [0.75, 1.0, 0.5], [1.0, 3.9, 1.0]
    coef = torch.rand(Cn, K, B, device="cuda", generator=g) - 0.5
    gl = torch.tensor([1.7], device="cuda")
    lpx_ref = torch.empty(Cn, K, B, device="cuda")
    g_ref = [torch.empty_like(r) for r in recons]
    for i in range(R):
        C.check(lib.mv_moe_lpx_fwd(C.ptr(recons[i]), C.dtype_code(recons[i]), C.ptr(xs[i]), C.ptr(lpx_ref), Cn, K, B, D, C.DIST[dist],
                                   scales[i], rescales[i], C.ptr(masks[i]), 1 if i else 0, C.stream()), "fwd")
        C.check(lib.mv_moe_lpx_bwd(C.ptr(recons[i]), C.dtype_code(recons[i]), C.ptr(xs[i]), C.ptr(coef), C.ptr(gl), C.ptr(g_ref[i]), Cn,
                                   K, B, D, C.DIST[dist], scales[i], rescales[i], C.ptr(masks[i]), C.stream()), "bwd")
    lpx = torch.full((Cn, K, B), 123.0, device="cuda")
    g_out = [torch.empty_like(r) for r in recons]
    C.check(lib.mv_moe_lpx_fwd_multi(R, C.ptr_array(recons), C.dtype_code(recons[0]), C.ptr_array(xs), C.ptr(lpx), Cn, K, B, D,
                                     C.DIST[dist], C.float_array(scales), C.float_array(rescales), C.ptr_array(masks), C.stream()),
            "fwd_multi")
    C.check(lib.mv_moe_lpx_bwd_multi(R, C.ptr_array(recons), C.dtype_code(recons[0]), C.ptr_array(xs), C.ptr(coef), C.ptr(gl),
                                     C.ptr_array(g_out), Cn, K, B, D, C.DIST[dist], C.float_array(scales), C.float_array(rescales),
                                     C.ptr_array(masks), C.stream()), "bwd_multi")
    torch.cuda.synchronize()
    assert float((lpx - lpx_ref).abs().max()) <= 1e-5 * float(lpx_ref.abs().max())
    for a, b in zip(g_out, g_ref):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("P,V", [(4, 5), (1, 40), (7, 100)])
def test_categorical_lpx_kernels_vs_torch(P, V, dtype):
    """mv_moe_lpx_cat_fwd / bwd against target * log_softmax(input + 1e-6) (base_utils.py:28-39)."""
    import torch.nn.functional as F
    from multivae_b200 import _cabi as C
    g = torch.Generator(device="cuda").manual_seed(3)
    Cn, K, B = 2, 3, 5
    recon = (torch.randn(Cn, K, B, P, V, device="cuda", generator=g) * 2).to(dtype)
    x = F.one_hot(torch.randint(0, V, (B, P), device="cuda", generator=g), V).float()
    mask = torch.tensor([1, 1, 0, 1, 1], dtype=torch.uint8, device="cuda")
    lpx = torch.empty(Cn, K, B, device="cuda")
    lib = C.lib()
    C.check(lib.mv_moe_lpx_cat_fwd(C.ptr(recon), C.dtype_code(recon), C.ptr(x), C.ptr(lpx), Cn, K, B, P, V, 1.3, C.ptr(mask), 0, C.stream()), "f")
    r32 = recon.float().requires_grad_(True)
    ref = (x * F.log_softmax(r32 + 1e-6, dim=-1)).mul(1.3).sum((-1, -2)) * mask.float()
    assert torch.allclose(lpx, ref, rtol=2e-5, atol=1e-4), float((lpx - ref).abs().max())
    coef = torch.randn(Cn, K, B, device="cuda", generator=g)
    gl = torch.tensor([0.7], device="cuda")
    gr = torch.empty_like(recon)
    C.check(lib.mv_moe_lpx_cat_bwd(C.ptr(recon), C.dtype_code(recon), C.ptr(x), C.ptr(coef), C.ptr(gl), C.ptr(gr), Cn, K, B, P, V, 1.3, C.ptr(mask), C.stream()), "b")
    (ref * coef).sum().mul(0.7).backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert torch.allclose(gr.float(), r32.grad, rtol=tol, atol=tol * float(r32.grad.abs().max()))


def test_logmeanexp_and_gauss_kl_kernels_vs_torch():
    import math
    from multivae_b200.elbo import kl_divergence, logmeanexp
    g = torch.Generator(device="cuda").manual_seed(5)
    lw = torch.randn(1000, 37, device="cuda", generator=g) * 50 - 12000
    ref = torch.logsumexp(lw.double(), 0) - math.log(1000)
    assert torch.allclose(logmeanexp(lw).double(), ref, rtol=1e-6, atol=1e-3)
    for prior_shape in [(1, 16), (9, 16)]:
        mu, lv = [torch.randn(9, 16, device="cuda", generator=g).requires_grad_(True) for _ in range(2)]
        pm, pl = [torch.randn(*prior_shape, device="cuda", generator=g).requires_grad_(True) for _ in range(2)]
        kl = kl_divergence(mu, lv, pm, pl)
        ref = (0.5 * (pl - lv + torch.exp(lv - pl) + (mu - pm) ** 2 / torch.exp(pl) - 1)).sum(-1)
        assert torch.allclose(kl, ref, rtol=1e-5, atol=1e-5)
        w = torch.randn(9, device="cuda", generator=g)
        got = torch.autograd.grad((kl * w).sum(), [mu, lv, pm, pl])
        want = torch.autograd.grad((ref * w).sum(), [mu, lv, pm, pl])
        for a, b in zip(got, want):
            assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-4, atol=1e-5)


def test_stable_poe_survives_extreme_log_variances():
    """stable_poe (base_utils.py:133-147) is a logsumexp: log-variances far outside exp()'s fp32 range must neither overflow
    nor underflow the product (MVAE path, prior expert always)."""
    from multivae_b200.elbo import poe_joint
    mu = torch.tensor([[[1.0, 2.0, -1.0]], [[3.0, -2.0, 0.5]]], device="cuda")           # (M=2, B=1, L=3)
    lv = torch.tensor([[[-200.0, 150.0, 0.0]], [[-190.0, 160.0, 0.0]]], device="cuda")
    bits = torch.tensor([3], dtype=torch.int32, device="cuda")
    jm, jl = poe_joint(mu, lv, None, bits, 1, True, 0.0)
    lv64, mu64 = torch.cat([lv.double(), torch.zeros(1, 1, 3, dtype=torch.float64, device="cuda")]), torch.cat([mu.double(), torch.zeros(1, 1, 3, dtype=torch.float64, device="cuda")])
    ref_lv = -torch.logsumexp(-lv64, 0)
    ref_mu = (torch.softmax(-lv64, 0) * mu64).sum(0)
    assert torch.isfinite(jm).all() and torch.isfinite(jl).all()
    assert torch.allclose(jl.double(), ref_lv, rtol=1e-6, atol=1e-6) and torch.allclose(jm.double(), ref_mu, rtol=1e-5, atol=1e-6)


SMALL_CASES = [n for n in sorted(CASES) if "arch" not in CASES[n] and not n.startswith("cfg")]


@pytest.mark.parametrize("name", SMALL_CASES)
def test_small_cases_on_the_native_networks(name):
    """Every small golden case (all models, losses, mask variants, categorical / Bernoulli decoders) with the encoders / decoders
    on the native tensor-core path (bf16 operands): the loss stays within 2e-2 of the fp32 reference (tiny MLPs, B = 5..9: the
    bf16 rounding of single samples does not average out), masked modalities still get exactly zero gradients, and every
    gradient is finite."""
    from tests.gpu_checks import rel, run_product
    out, model, rec = run_product(name, compute_dtype=torch.bfloat16)
    assert rel(out.loss.detach().cpu(), rec["loss"]) <= 2e-2, (name, float(out.loss), float(rec["loss"]))
    for k, p in model.named_parameters():
        g = rec["grads"][k]
        if g is None:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, (name, k, "expected no gradient")
        else:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), (name, k)
