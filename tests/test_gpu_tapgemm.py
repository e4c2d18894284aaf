"""tcgen05 tap-GEMM (mv_tapgemm) against plain PyTorch fp32 references of the same ops on the same
bf16-rounded operands: 3x3 / 1x1 convolutions in the shared-halo layout (forward and data-gradient
packs), Linear, the fused epilogues and the NCHW image head."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.rand(*shape, device="cuda", generator=g) * 2 - 1) * scale


def _close(got, ref, tol=2e-2):
    err = float((got.float() - ref).abs().max())
    den = float(ref.abs().max()) + 1e-6
    assert err / den <= tol, (err, den)


@pytest.mark.parametrize("n_img,H,cin,cout", [(37, 28, 64, 64), (300, 28, 64, 64), (23, 14, 128, 64), (41, 7, 256, 128),
                                              (19, 7, 128, 128), (9, 14, 64, 64),
                                              # the CUB ResNet stages; 64-pixel rows: the input window spans two TMA boxes
                                              (5, 64, 64, 64), (3, 32, 128, 64), (3, 32, 64, 128), (3, 16, 256, 128), (2, 16, 128, 256)])
def test_conv3x3_forward_epilogues(n_img, H, cin, cout):
    from multivae_b200.nn import halo as HL
    x = _rnd(n_img, cin, H, H, seed=1).bfloat16()
    w = _rnd(cout, cin, 3, 3, seed=2, scale=cin ** -0.5).bfloat16()
    b = _rnd(cout, seed=3)
    r = _rnd(n_img, cout, H, H, seed=4).bfloat16()
    A, g = HL.to_halo(x)
    R, _ = HL.to_halo(r)
    act = torch.empty(g.P, cout, device="cuda", dtype=torch.bfloat16)
    out = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), cout, g.P, bias=b, act="lrelu", alpha=0.1, res=R,
                     out2=act, out2_pre=True, geom=g)
    y = F.leaky_relu(F.conv2d(x.float(), w.float(), b, padding=1), 0.2)
    _close(HL.from_halo(act, g), y)
    _close(HL.from_halo(out, g), r.float() + 0.1 * y)
    # halo rows must stay exactly zero (they are the padding of the next convolution)
    mask = torch.ones(g.P, dtype=torch.bool, device="cuda")
    v = mask[: n_img * g.S].view(n_img, H + 1, g.Wp)
    v[:, 1:, :H] = False
    assert float(out[mask].abs().max()) == 0.0 and float(act[mask].abs().max()) == 0.0


@pytest.mark.parametrize("n_img,H,cin,cout", [(21, 28, 64, 64), (17, 14, 128, 64), (11, 7, 256, 128), (4, 64, 64, 64), (3, 32, 128, 64),
                                              (3, 16, 256, 128)])
def test_conv3x3_dgrad_with_activation_derivative(n_img, H, cin, cout):
    from multivae_b200.nn import halo as HL
    gy = _rnd(n_img, cout, H, H, seed=5).bfloat16()
    w = _rnd(cout, cin, 3, 3, seed=6, scale=cout ** -0.5).bfloat16()
    saved = _rnd(n_img, cin, H, H, seed=7).bfloat16()   # activation of the layer below (sign decides lrelu')
    short = _rnd(n_img, cin, H, H, seed=8).bfloat16()
    G, g = HL.to_halo(gy)
    S_, _ = HL.to_halo(saved)
    R, _ = HL.to_halo(short)
    out = HL.tapgemm(G, HL.pack_conv_weight_dgrad(w), 9, g.taps3x3(), cin, g.P, dact1=S_, res=R, geom=g)
    ref = F.conv_transpose2d(gy.float(), w.float(), padding=1) * torch.where(saved.float() > 0, 1.0, 0.2) + short.float()
    _close(HL.from_halo(out, g), ref)


@pytest.mark.parametrize("n_img,H,cin", [(37, 28, 64), (301, 28, 64), (23, 14, 128), (9, 14, 64), (5, 7, 64), (4, 64, 64), (3, 32, 128)])
@pytest.mark.parametrize("variant", ["bias_lrelu", "dact1", "res", "none"])
def test_conv3x3_cout64_single_side_variants(n_img, H, cin, variant):
    """The three-taps-per-MMA kernel (csrc/tapconv3.cu) behind mv_tapgemm for 3x3 convolutions with 64 outputs."""
    from multivae_b200.nn import halo as HL
    cout = 64
    x = _rnd(n_img, cin, H, H, seed=21).bfloat16()
    w = _rnd(cout, cin, 3, 3, seed=22, scale=cin ** -0.5).bfloat16()
    b = _rnd(cout, seed=23)
    s = _rnd(n_img, cout, H, H, seed=24).bfloat16()
    A, g = HL.to_halo(x)
    S_, _ = HL.to_halo(s)
    conv = F.conv2d(x.float(), w.float(), padding=1)
    if variant == "bias_lrelu":
        out = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), cout, g.P, bias=b, act="lrelu", geom=g)
        ref = F.leaky_relu(conv + b.view(1, -1, 1, 1), 0.2)
    elif variant == "dact1":
        out = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), cout, g.P, dact1=S_, geom=g)
        ref = conv * torch.where(s.float() > 0, 1.0, 0.2)
    elif variant == "res":
        out = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), cout, g.P, res=S_, geom=g)
        ref = conv + s.float()
    else:
        out = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), cout, g.P, geom=g)
        ref = conv
    _close(HL.from_halo(out, g), ref)
    mask = torch.ones(g.P, dtype=torch.bool, device="cuda")
    v = mask[: n_img * g.S].view(n_img, H + 1, g.Wp)
    v[:, 1:, :H] = False
    assert float(out[mask].abs().max()) == 0.0


def test_conv1x1_and_two_outputs():
    from multivae_b200.nn import halo as HL
    n_img, H, cin, cout = 29, 14, 128, 64
    x = _rnd(n_img, cin, H, H, seed=9).bfloat16()
    w = _rnd(cout, cin, 1, 1, seed=10, scale=cin ** -0.5).bfloat16()
    d = _rnd(n_img, cout, H, H, seed=11).bfloat16()
    A, g = HL.to_halo(x)
    D, _ = HL.to_halo(d)
    o2 = torch.empty(g.P, cout, device="cuda", dtype=torch.bfloat16)
    out = HL.tapgemm(A, HL.pack_conv_weight(w), 1, [0], cout, g.P, out2=o2, alpha2=0.1, dact2=D, geom=g)
    ref = F.conv2d(x.float(), w.float())
    _close(HL.from_halo(out, g), ref)
    _close(HL.from_halo(o2, g), 0.1 * ref * torch.where(d.float() > 0, 1.0, 0.2))


@pytest.mark.parametrize("n_img,H", [(33, 28), (3, 64)])
def test_image_head_nchw_and_cin16(n_img, H):
    from multivae_b200.nn import halo as HL
    x = _rnd(n_img, 64, H, H, seed=12).bfloat16()
    w = _rnd(3, 64, 3, 3, seed=13, scale=0.05).bfloat16()
    b = _rnd(3, seed=14)
    A, g = HL.to_halo(x)
    wp = torch.zeros(16, 64, 3, 3, device="cuda", dtype=torch.bfloat16)
    wp[:3] = w
    bp = torch.zeros(16, device="cuda")
    bp[:3] = b
    img = torch.empty(n_img, 3, H, H, device="cuda", dtype=torch.bfloat16)
    HL.tapgemm(A, HL.pack_conv_weight(wp), 9, g.taps3x3(), 16, g.P, bias=bp, act="lrelu", geom=g, nchw_out=img, n_valid=3)
    _close(img, F.leaky_relu(F.conv2d(x.float(), w.float(), b, padding=1), 0.2))
    # data gradient of the head: 16 (3 real) channels in, 64 out  -> the SWIZZLE_32B operand path
    gy = torch.zeros(n_img, 16, H, H, device="cuda", dtype=torch.bfloat16)
    gy[:, :3] = _rnd(n_img, 3, H, H, seed=15).bfloat16()
    G, _ = HL.to_halo(gy)
    out = HL.tapgemm(G, HL.pack_conv_weight_dgrad(wp), 9, g.taps3x3(), 64, g.P, geom=g)
    _close(HL.from_halo(out, g), F.conv_transpose2d(gy[:, :3].float(), w.float(), padding=1))


@pytest.mark.parametrize("rows,K,N", [(1000, 64, 16384), (333, 512, 256), (128, 16384, 64)])
def test_linear(rows, K, N):
    from multivae_b200.nn import halo as HL
    x = _rnd(rows, K, seed=16).bfloat16()
    w = _rnd(N, K, seed=17, scale=K ** -0.5).bfloat16()
    b = _rnd(N, seed=18)
    out = HL.tapgemm(x, w, 1, [0], N, rows, bias=b, act="relu")
    _close(out, F.relu(F.linear(x.float(), w.float(), b)))


def test_pack_conv_weights_kernel_matches_torch_packs():
    """mv_pack_conv_weights (all packs of a network in one launch) against pack_conv_weight / pack_conv_weight_dgrad."""
    from multivae_b200.nn import halo as HL
    ws = [_rnd(64, 64, 3, 3, seed=31), _rnd(128, 256, 1, 1, seed=32), _rnd(3, 64, 3, 3, seed=33), _rnd(64, 3, 3, 3, seed=34)]
    specs = [(ws[0], 64, 64, True), (ws[1], 128, 256, True), (ws[2], 16, 64, True), (ws[3], 64, 16, False)]
    packs = HL.pack_conv_weights(specs)
    for (w, npad, cpad, want_d), (fwd, dg) in zip(specs, packs):
        wp = torch.zeros(npad, cpad, *w.shape[2:], device="cuda")
        wp[: w.shape[0], : w.shape[1]] = w
        assert torch.equal(fwd, HL.pack_conv_weight(wp))
        if want_d:
            assert torch.equal(dg, HL.pack_conv_weight_dgrad(wp))
        else:
            assert dg is None


@pytest.mark.parametrize("n_img", [37, 301])
def test_sign_mask_second_output_and_its_consumer(n_img):
    """out2_mask (conv3 kernel: the saved activation as one sign bit per element) and dmask2 (tap-GEMM epilogue reading it)
    against the bf16 `out2` / `dact2` path: identical first outputs, mask == (out2 > 0), identical consumer outputs."""
    from multivae_b200.nn import halo as HL
    H, c = 28, 64
    x = _rnd(n_img, c, H, H, seed=41).bfloat16()
    w = _rnd(c, c, 3, 3, seed=42, scale=c ** -0.5).bfloat16()
    b = _rnd(c, seed=43)
    r = _rnd(n_img, c, H, H, seed=44).bfloat16()
    A, g = HL.to_halo(x)
    R, _ = HL.to_halo(r)
    d = torch.empty(g.P, c, device="cuda", dtype=torch.bfloat16)
    out_ref = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), c, g.P, bias=b, act="lrelu", alpha=0.1, res=R, out2=d, out2_pre=True, geom=g)
    mask = torch.zeros(HL.mask_rows(g.P), device="cuda", dtype=torch.int64)
    out = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), c, g.P, bias=b, act="lrelu", alpha=0.1, res=R, out2_mask=mask, geom=g)
    assert torch.equal(out, out_ref)
    bits = ((mask[: g.P].unsqueeze(1) >> torch.arange(64, device="cuda")) & 1).bool()
    assert torch.equal(bits, d > 0)
    # consumer: 16 -> 64 channel data gradient with a second output scaled by lrelu'(d)
    gy = torch.zeros(n_img, 16, H, H, device="cuda", dtype=torch.bfloat16)
    gy[:, :3] = _rnd(n_img, 3, H, H, seed=45).bfloat16()
    G, _ = HL.to_halo(gy)
    wd = HL.pack_conv_weight_dgrad(torch.cat([_rnd(3, c, 3, 3, seed=46, scale=0.05), torch.zeros(13, c, 3, 3, device="cuda")]).bfloat16())
    o2a = torch.empty(g.P, c, device="cuda", dtype=torch.bfloat16)
    o2b = torch.empty(g.P, c, device="cuda", dtype=torch.bfloat16)
    oa = HL.tapgemm(G, wd, 9, g.taps3x3(), c, g.P, out2=o2a, alpha2=0.1, dact2=d, geom=g)
    ob = HL.tapgemm(G, wd, 9, g.taps3x3(), c, g.P, out2=o2b, alpha2=0.1, dmask2=mask, geom=g)
    assert torch.equal(oa, ob) and torch.equal(o2a, o2b)
    # single-output form of the same consumer: out = 0.1 * acc * lrelu'(d) from the sign bits (dmask1) ...
    oc = HL.tapgemm(G, wd, 9, g.taps3x3(), c, g.P, alpha=0.1, dmask1=mask, slope1=0.2, geom=g)
    _close(oc, o2a.float(), tol=1e-2)
    # ... and a 3x3 convolution whose residual is recovered from it: res * (bit ? 10 : 50) = the un-scaled gradient `oa`
    gh = _rnd(n_img, c, H, H, seed=47).bfloat16()
    GH, _ = HL.to_halo(gh)
    w2 = HL.pack_conv_weight_dgrad(_rnd(c, c, 3, 3, seed=48, scale=c ** -0.5).bfloat16())
    ref = HL.tapgemm(GH, w2, 9, g.taps3x3(), c, g.P, res=oa, geom=g)
    got = HL.tapgemm(GH, w2, 9, g.taps3x3(), c, g.P, res=oc, res_mask=mask, res_scale=(10.0, 50.0), geom=g)
    _close(got, ref.float(), tol=1e-2)
    halo = torch.ones(g.P, dtype=torch.bool, device="cuda")
    v = halo[: n_img * g.S].view(n_img, H + 1, g.Wp)
    v[:, 1:, :H] = False
    assert float(got[halo].abs().max()) == 0.0 and float(oc[halo].abs().max()) == 0.0


@pytest.mark.parametrize("n_img,H", [(23, 14), (301, 14), (7, 28)])
def test_conv3x3_with_fused_1x1_term(n_img, H):
    """out = conv3x3(x; w) + a2 @ w2^T in one launch (the conv0 data gradient of a ResnetBlock with a learned shortcut)."""
    from multivae_b200.nn import halo as HL
    x = _rnd(n_img, 64, H, H, seed=61).bfloat16()
    w = _rnd(64, 64, 3, 3, seed=62, scale=64 ** -0.5).bfloat16()
    x2 = _rnd(n_img, 64, H, H, seed=63).bfloat16()
    w2 = _rnd(64, 64, seed=64, scale=0.125).bfloat16()
    A, g = HL.to_halo(x)
    A2, _ = HL.to_halo(x2)
    out = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), 64, g.P, a2=A2, w2=w2.contiguous(), geom=g)
    ref = F.conv2d(x.float(), w.float(), padding=1) + F.conv2d(x2.float(), w2.float().view(64, 64, 1, 1))
    _close(HL.from_halo(out, g), ref)
    mask = torch.ones(g.P, dtype=torch.bool, device="cuda")
    v = mask[: n_img * g.S].view(n_img, H + 1, g.Wp)
    v[:, 1:, :H] = False
    assert float(out[mask].abs().max()) == 0.0
    # into a column slice of a wider tensor (how the 128-channel gradient of the block is assembled)
    wide = torch.zeros(g.P, 128, device="cuda", dtype=torch.bfloat16)
    HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), 64, g.P, a2=A2, w2=w2.contiguous(), out=wide[:, 64:], geom=g)
    assert torch.equal(wide[:, 64:], out) and float(wide[:, :64].abs().max()) == 0.0
