"""Strided / transposed convolution path (csrc/im2col.cu + mv_gemm): gather kernels against F.unfold / F.fold, and the SVHN /
conv-PolyMNIST networks on the native path against goldens of the REAL reference modules."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _r(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.rand(*shape, device="cuda", generator=g) * 2 - 1


@pytest.mark.parametrize("C,H,k,s,p", [(3, 32, 4, 2, 1), (32, 16, 4, 2, 1), (3, 28, 3, 2, 1), (64, 7, 3, 2, 1), (128, 4, 4, 1, 0)])
@pytest.mark.parametrize("nchw", [True, False])
@pytest.mark.parametrize("tc", [False, True])
def test_im2col_matches_unfold(C, H, k, s, p, nchw, tc):
    from multivae_b200.nn.conv_native import _geom, _pad8, im2col
    n = 5
    x = _r(n, C, H, H, seed=1)
    Ho = (H + 2 * p - k) // s + 1
    src = x.contiguous() if nchw else x.permute(0, 2, 3, 1).contiguous().bfloat16()
    g = _geom(n, H, H, C, nchw, k, s, p, Ho, Ho, _pad8(C * k * k), tc=tc)
    cols = im2col(src, g)
    ref = F.unfold(x if nchw else src.permute(0, 3, 1, 2).float(), k, stride=s, padding=p).transpose(1, 2).reshape(n * Ho * Ho, C * k * k)
    if tc:   # tap-major columns: (c, t) -> (t, c)
        ref = ref.view(-1, C, k * k).transpose(1, 2).reshape(-1, C * k * k)
    assert torch.equal(cols[:, : C * k * k].float(), ref.bfloat16().float())
    assert float(cols[:, C * k * k:].abs().sum()) == 0.0


@pytest.mark.parametrize("C,Hi,k,s,p,op", [(64, 4, 4, 2, 1, 0), (3, 16, 4, 2, 1, 0), (64, 4, 3, 2, 1, 0), (32, 7, 3, 2, 1, 1), (3, 14, 3, 2, 1, 1), (128, 1, 4, 1, 0, 0)])
@pytest.mark.parametrize("nchw", [True, False])
@pytest.mark.parametrize("tc", [False, True])
def test_col2im_matches_fold(C, Hi, k, s, p, op, nchw, tc):
    from multivae_b200.nn.conv_native import _geom, _pad8, col2im
    n = 4
    Ho = (Hi - 1) * s - 2 * p + k + op
    T = k * k
    ld = _pad8(C * T)
    cols = torch.zeros(n * Hi * Hi, ld, device="cuda", dtype=torch.bfloat16)
    cols[:, : C * T] = _r(n * Hi * Hi, C * T, seed=2).bfloat16()
    bias = _r(C, seed=3)
    g = _geom(n, Ho, Ho, C, nchw, k, s, p, Hi, Hi, ld, tc=tc)
    cols_ct = cols
    if tc:   # hand the kernel the tap-major version of the same patch matrix
        cols = torch.zeros_like(cols_ct)
        cols[:, : C * T] = cols_ct[:, : C * T].view(-1, C, T).transpose(1, 2).reshape(-1, C * T)
    out = col2im(cols, g, bias=bias, act="relu")
    out32 = col2im(cols.float(), g, bias=bias, act="relu")   # fp32 patch matrix (what the GEMMs hand over)
    assert float((out32.float() - out.float()).abs().max()) <= 1e-2 * max(1.0, float(out.float().abs().max()))
    ref = F.fold(cols_ct[:, : C * T].float().reshape(n, Hi * Hi, C * T).transpose(1, 2), (Ho, Ho), k, stride=s, padding=p) if op == 0 else None
    if ref is None:   # output_padding: fold onto the padded canvas by hand
        full = F.fold(cols_ct[:, : C * T].float().reshape(n, Hi * Hi, C * T).transpose(1, 2), (Ho + p, Ho + p), k, stride=s, padding=0)
        ref = full[:, :, p:p + Ho, p:p + Ho]
    ref = torch.relu(ref + bias.view(1, C, 1, 1))
    got = out.float() if nchw else out.float().view(n, Ho, Ho, C).permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2e-2 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("net", ["enc_svhn", "dec_svhn", "enc_conv_mmnist", "dec_conv_mmnist"])
def test_native_conv_networks_match_reference_goldens(net):
    from tests.test_gpu_gemm import native_vs_library_bf16
    e_nat, e_lib = native_vs_library_bf16(net)
    assert e_nat["out"] <= 3e-2
    # goldens of 2-3 samples: two bf16 implementations with different rounding points scatter by tens of percent around each other
    for k in ("out", "grad", "grad_in", "grad_in_l2"):
        assert e_nat[k] <= 1.5 * e_lib[k] + 5e-3, (net, k, e_nat[k], e_lib[k])


def test_pack_tc_and_unpack_tc_add_roundtrip():
    from multivae_b200.nn.conv_native import pack_tc, unpack_tc_add
    ws = [_r(32, 3, 4, 4, seed=5), _r(64, 32, 3, 3, seed=6), _r(20, 128, 4, 4, seed=7)]
    packs = pack_tc(ws)
    for w, p in zip(ws, packs):
        n, c, kh, kw = w.shape
        ref = w.permute(0, 2, 3, 1).reshape(n, kh * kw * c)
        assert torch.equal(p[:, : ref.shape[1]].float(), ref.bfloat16().float())
        assert float(p[:, ref.shape[1]:].abs().sum()) == 0.0
    grads = [torch.ones_like(w) for w in ws]
    dWs = [_r(w.shape[0], w[0].numel(), seed=8 + i) for i, w in enumerate(ws)]
    unpack_tc_add(list(zip(dWs, grads)))
    for w, d, g in zip(ws, dWs, grads):
        n, c, kh, kw = w.shape
        assert torch.allclose(g, 1 + d.view(n, kh * kw, c).transpose(1, 2).reshape(n, c, kh, kw))
