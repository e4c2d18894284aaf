"""mv_gemm (general tcgen05 GEMM, K-major / MN-major operands, three output kinds) and the native MLP chain against torch fp32
on the same bf16-rounded operands, including ragged sizes (no padding of M, N, K) and the BASELINE configs' Linear shapes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return ((torch.rand(*shape, device="cuda", generator=g) * 2 - 1) * scale)


def _pitch8(t):
    """bf16 copy with a row pitch that is a multiple of 8 elements, as a [rows, cols] view."""
    from multivae_b200.nn.linear_native import _to_bf16_padded
    return _to_bf16_padded(t)


SHAPES = [(256, 512, 784), (32, 784, 512), (2560, 40, 512), (200, 512, 20), (77, 136, 72), (640, 3072, 512), (130, 12288, 512), (129, 128, 12288)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_forward_gemm_bias_act(M, N, K):
    from multivae_b200.nn.linear_native import _pad8, gemm
    X, W, b = _pitch8(_r(M, K, seed=1)), _pitch8(_r(N, K, seed=2, scale=K ** -0.5)), _r(N, seed=3)
    ref = X.float() @ W.float().t() + b
    for act, f in (("none", lambda t: t), ("relu", torch.relu), ("sigmoid", torch.sigmoid)):
        out = torch.full((M, _pad8(N)), 7.0, device="cuda", dtype=torch.bfloat16)[:, :N]
        gemm(X, W, M, N, K, out, bias=b, act=act)
        want = f(ref)
        assert float((out.float() - want).abs().max()) <= 1e-2 * max(1.0, float(want.abs().max())), (act, float((out.float() - want).abs().max()))
    o32 = torch.full((M, N), 7.0, device="cuda")
    gemm(X, W, M, N, K, o32, bias=b, out_kind=1)
    assert float((o32 - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_data_gradient_gemm_mn_major_b_with_relu_mask(M, N, K):
    """dX[b, k] = sum_n dY[b, n] W[n, k] (* relu'(h[b, k])): B = W stored [n][k] read MN-major."""
    from multivae_b200.nn.linear_native import _pad8, gemm
    dY, W = _pitch8(_r(M, N, seed=4)), _pitch8(_r(N, K, seed=5, scale=N ** -0.5))
    ref = dY.float() @ W.float()
    o32 = torch.empty(M, K, device="cuda")
    gemm(dY, W, M, K, N, o32, b_mn=True, out_kind=1)
    assert float((o32 - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))
    if K % 8 == 0:
        h = _pitch8(_r(M, K, seed=6))
        out = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
        gemm(dY, W, M, K, N, out, b_mn=True, dact=h, dslope=0.0)
        want = ref * (h.float() > 0)
        assert float((out.float() - want).abs().max()) <= 1e-2 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_weight_gradient_gemm_both_mn_major_accumulates(M, N, K):
    """dW[n, k] += sum_b dY[b, n] X[b, k]: both operands MN-major, fp32 reductions (split over the reduction when few tiles)."""
    from multivae_b200.nn.linear_native import gemm
    dY, X = _pitch8(_r(M, N, seed=7)), _pitch8(_r(M, K, seed=8))
    ref = dY.float().t() @ X.float()
    dW = torch.ones(N, K, device="cuda")
    gemm(dY, X, N, K, M, dW, a_mn=True, b_mn=True, out_kind=2)
    assert float((dW - 1 - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), float((dW - 1 - ref).abs().max())


def test_colsum_any_and_act_bwd():
    from multivae_b200 import _cabi as C
    lib = C.lib()
    for P, N in [(2560, 12288), (300, 40), (64, 784)]:
        G = _r(P, N, seed=9).bfloat16()
        out = torch.ones(N, device="cuda")
        C.check(lib.mv_colsum_any(G.data_ptr(), P, N, N, out.data_ptr(), C.stream()), "colsum")
        ref = G.float().sum(0) + 1
        assert float((out - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
    for dt in (torch.float32, torch.bfloat16):
        g, y = _r(50, 784, seed=10).to(dt), torch.sigmoid(_r(50, 784, seed=11) * 3).bfloat16()
        o = torch.empty(50, 784, device="cuda", dtype=torch.bfloat16)
        C.check(lib.mv_act_bwd(C.ptr(g), C.dtype_code(g), C.ptr(y), C.ptr(o), 50 * 784, 3, 0.0, C.stream()), "act")
        want = g.float() * y.float() * (1 - y.float())
        assert float((o.float() - want).abs().max()) <= 1e-2 * float(want.abs().max())


def native_vs_library_bf16(net):
    """Errors of a module against the REAL reference's golden on (a) the native tensor-core path and (b) the same module run by
    the library under torch's bf16 autocast: the native path must be at least as accurate as the library's own bf16 arithmetic
    (the remaining distance to the fp32 reference is the operand precision BASELINE.json asks for, not the kernels)."""
    from multivae_b200.nn import functional as NF
    from tests.net_checks import check_net
    NF.set_backend("native")
    try:
        e_nat = check_net(net, "cuda", collect=True, verbose=True)
    finally:
        NF.set_backend("torch")
    try:
        e_lib = check_net(net, "cuda", collect=True, autocast=True, verbose=True)
    finally:
        NF.set_backend("auto")
    return e_nat, e_lib


@pytest.mark.parametrize("net", ["enc_mlp", "enc_mlp_style", "dec_mlp"])
def test_native_mlp_modules_match_reference_goldens(net):
    e_nat, e_lib = native_vs_library_bf16(net)
    assert e_nat["out"] <= 2e-2
    for k in ("out", "grad", "grad_in", "grad_in_l2"):
        assert e_nat[k] <= 1.25 * e_lib[k] + 2e-3, (net, k, e_nat[k], e_lib[k])


def test_native_mlp_chain_gradients_vs_torch_fp32_large():
    """A cfg5-sized decoder (64 -> 512 -> 12288, sigmoid) on 2560 rows: outputs and all gradients vs torch fp32 on the same
    weights, bounded by the error of torch's own bf16 autocast of the same modules."""
    import torch.nn as nn
    from multivae_b200.nn.linear_native import mlp_chain
    torch.manual_seed(0)
    l0, l1 = nn.Linear(64, 512).cuda(), nn.Linear(512, 12288).cuda()
    z = _r(2560, 64, seed=12).requires_grad_(True)
    cot = _r(2560, 12288, seed=13)
    params = [z] + list(l0.parameters()) + list(l1.parameters())

    def grads(y):
        gs = torch.autograd.grad((y.float() * cot).sum(), params)
        return [g.clone() for g in gs]

    y = mlp_chain(z, [l0, l1], ["relu", "sigmoid"])
    got = grads(y)
    yr = torch.sigmoid(l1(torch.relu(l0(z))))
    want = grads(yr)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = torch.sigmoid(l1(torch.relu(l0(z))))
    auto = grads(ya)
    assert float((y.float() - yr).abs().max()) <= 1e-2
    for a, c, b, name in zip(got, auto, want, ["z", "w0", "b0", "w1", "b1"]):
        rel, rel_lib = float((a - b).norm() / b.norm()), float((c - b).norm() / b.norm())
        assert rel <= 1.25 * rel_lib + 1e-3, (name, rel, rel_lib)


def test_gather_kernels_vs_torch_indexing():
    """mv_gather_cast / mv_gather_f32 (the halo-order layout of the fc weights and its inverse on the gradients): row and column
    gathers with out-of-range indices reading as zeros, padded row pitch, accumulation."""
    from multivae_b200.nn.resnet_native import _gather_cast, _gather_f32
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randn(37, 21, device="cuda", generator=g)
    idx_r = torch.randint(0, 38, (50,), device="cuda", generator=g)           # 37 = "zero row"
    ext = torch.cat([src, src.new_zeros(1, 21)], 0)
    got = _gather_cast(src, idx_r, 0, 50, 21)
    assert got.shape == (50, 21) and got.stride(0) == 24
    assert torch.equal(got, ext[idx_r].bfloat16())
    idx_c = torch.randint(0, 22, (30,), device="cuda", generator=g)           # 21 = "zero column"
    extc = torch.cat([src, src.new_zeros(37, 1)], 1)
    assert torch.equal(_gather_cast(src, idx_c, 1, 37, 30), extc[:, idx_c].bfloat16())
    assert torch.equal(_gather_f32(src, 21, idx_r, 0, 50, 21), ext[idx_r])
    assert torch.equal(_gather_f32(src, 21, idx_c, 1, 37, 30), extc[:, idx_c])
    base = torch.randn(50, 21, device="cuda", generator=g)
    acc = base.clone()
    _gather_f32(src, 21, idx_r, 0, 50, 21, into=acc)
    assert torch.allclose(acc, base + ext[idx_r])
    v = torch.randn(37, device="cuda", generator=g)
    assert torch.equal(_gather_f32(v, 1, idx_r.clamp(max=36), 0, 50, 1).view(50), v[idx_r.clamp(max=36)])
