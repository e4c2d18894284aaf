"""Size-independent properties of the hot-path kernels at the NORTH-STAR sizes (M*K*B = 12800 images per decoder at 28x28,
C=5, K=10, B=256, D=2352 for the ELBO kernels), where the oracle cannot be run: exact linearity under scaling by 2, exact
independence of an image's outputs from the other images in the launch, additivity of the weight gradient over image
subsets, zero halo rows."""
import pytest
import torch

pytestmark = pytest.mark.gpu

N_IMG = 12800


def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.rand(*shape, device="cuda", generator=g) * 2 - 1) * scale


def _halo_mask(g):
    mask = torch.ones(g.P, dtype=torch.bool, device="cuda")
    v = mask[: g.n_img * g.S].view(g.n_img, g.H + 1, g.Wp)
    v[:, 1:, : g.W] = False
    return mask


@pytest.mark.parametrize("variant", ["plain", "res"])
def test_conv3x3_full_size_linearity_and_image_independence(variant):
    from multivae_b200.nn import halo as HL
    g = HL.Geom(N_IMG, 28, 28)
    A = _rnd(g.P, 64, seed=1).bfloat16()
    A[_halo_mask(g)] = 0
    W = _rnd(9 * 64, 64, seed=2, scale=0.05).bfloat16()
    R = _rnd(g.P, 64, seed=3).bfloat16() if variant == "res" else None
    out = HL.tapgemm(A, W, 9, g.taps3x3(), 64, g.P, res=R, geom=g)
    # scaling every input by 2 is exact in bf16 / fp32: the output must be exactly doubled
    out2 = HL.tapgemm(A * 2, W, 9, g.taps3x3(), 64, g.P, res=None if R is None else R * 2, geom=g)
    assert torch.equal(out2.float(), out.float() * 2)
    # the first 50 images of the big launch == a launch on those 50 images alone (same tiling from row 0)
    gs = HL.Geom(50, 28, 28)
    sub = HL.tapgemm(A[: gs.P].clone(), W, 9, gs.taps3x3(), 64, gs.P, res=None if R is None else R[: gs.P].clone(), geom=gs)
    assert torch.equal(sub[: 50 * gs.S], out[: 50 * gs.S])
    # halo rows stay exactly zero
    assert float(out[_halo_mask(g)].abs().max()) == 0.0


def test_wgrad_full_size_additivity_over_images():
    from multivae_b200.nn import halo as HL
    g = HL.Geom(N_IMG, 28, 28)
    X = _rnd(g.P, 64, seed=4).bfloat16()
    G = _rnd(g.P, 64, seed=5, scale=0.1).bfloat16()
    m = _halo_mask(g)
    X[m] = 0
    G[m] = 0
    dW, db = HL.wgrad(X, G, 9, g.taps3x3(), g.P, want_db=True)
    # split the images in two launches (cut on an image boundary): gradients add up
    h = N_IMG // 2
    gh = HL.Geom(h, 28, 28)
    cut = h * g.S
    Xa, Ga = X[: gh.P].clone(), G[: gh.P].clone()
    Xa[cut:] = 0
    Ga[cut:] = 0
    dWa, dba = HL.wgrad(Xa, Ga, 9, gh.taps3x3(), gh.P, want_db=True)
    Xb = torch.cat([X[cut:], X.new_zeros(0, 64)])
    Gb = G[cut:]
    gb = HL.Geom(N_IMG - h, 28, 28)
    dWb, dbb = HL.wgrad(Xb.contiguous(), Gb.contiguous(), 9, gb.taps3x3(), gb.P, want_db=True)
    ref = dWa + dWb
    assert float((dW - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert float((db - (dba + dbb)).abs().max()) <= 1e-4 * float(db.abs().max()) + 1e-3


def test_lpx_full_size_rows_are_independent():
    """mv_moe_lpx_fwd_multi / bwd_multi on the north-star tensor sizes: the result for conditioning modality c does not
    depend on the other c (bit-exact), and doubling the upstream coefficient doubles the gradient exactly."""
    from multivae_b200 import _cabi as C
    Cn, K, B, D, R = 5, 10, 256, 2352, 5
    lib = C.lib()
    recons = [_rnd(Cn, K, B, D, seed=10 + i).bfloat16() for i in range(R)]
    xs = [_rnd(B, D, seed=20 + i).abs() for i in range(R)]
    sc, rs, none = C.float_array([0.75] * R), C.float_array([1.0] * R), C.ptr_array([None] * R)
    lpx = torch.empty(Cn, K, B, device="cuda")
    C.check(lib.mv_moe_lpx_fwd_multi(R, C.ptr_array(recons), 1, C.ptr_array(xs), C.ptr(lpx), Cn, K, B, D, 1, sc, rs, none, C.stream()), "f")
    one = [r[2:3].contiguous() for r in recons]
    lpx1 = torch.empty(1, K, B, device="cuda")
    C.check(lib.mv_moe_lpx_fwd_multi(R, C.ptr_array(one), 1, C.ptr_array(xs), C.ptr(lpx1), 1, K, B, D, 1, sc, rs, none, C.stream()), "f1")
    torch.cuda.synchronize()
    assert float((lpx[2] - lpx1[0]).abs().max()) <= 2e-6 * float(lpx1.abs().max())   # fp32 atomics over the 5 modalities
    assert torch.isfinite(lpx).all()
    coef = _rnd(Cn, K, B, seed=30)
    gl = torch.ones(1, device="cuda")
    g1 = [torch.empty_like(r) for r in recons]
    g2 = [torch.empty_like(r) for r in recons]
    C.check(lib.mv_moe_lpx_bwd_multi(R, C.ptr_array(recons), 1, C.ptr_array(xs), C.ptr(coef), C.ptr(gl), C.ptr_array(g1), Cn, K, B, D, 1,
                                     sc, rs, none, C.stream()), "b")
    C.check(lib.mv_moe_lpx_bwd_multi(R, C.ptr_array(recons), 1, C.ptr_array(xs), C.ptr(coef * 2), C.ptr(gl), C.ptr_array(g2), Cn, K, B, D,
                                     1, sc, rs, none, C.stream()), "b2")
    torch.cuda.synchronize()
    for a, b in zip(g1, g2):
        assert torch.equal(b.float(), a.float() * 2)
