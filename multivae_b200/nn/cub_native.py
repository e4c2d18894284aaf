"""Native (tcgen05) execution of the 64 x 64 CUB ResNets (reference: models/nn/cub.py:144-293).

Same kernels and layout as the PolyMNIST ResNets (resnet_native.py): activations in the shared-halo NHWC bf16 layout, 3x3 / 1x1
convolutions and their data gradients on mv_tapgemm, weight gradients on mv_wgrad, pooling / up-sampling / packing on the
HBM-bound helpers, the fully connected layers on mv_gemm.  The 64-pixel-wide stages need input windows of 128 + 2 * 66 = 260 rows,
more than one TMA box (256 rows): the kernels load them as two boxes.

The blocks are pre-activation blocks, `out = x_s + 0.1 * conv_1(actvn(conv_0(actvn(x))))` (cub.py:294-300): the shortcut reads
the raw x, conv_0 reads a = actvn(x) (mv_lrelu_fwd).  Backward of one block, g = dL/d out:
    dW1 = 0.1 * sum g h,   g_h = 0.1 * lrelu'(h) * conv_1^T(g)      (h = actvn(conv_0(a)), saved)
    dW0 = sum g_h a,       g_x = shortcut^T(g) + lrelu'(a) * conv_0^T(g_h)   (one mv_tapgemm launch with two side inputs)
There is no fallback inside this path: if the C-ABI library is missing, it raises."""
import torch

from .. import _cabi as C
from ..containers import ModelOutput
from . import halo as HL
from . import resnet_native as RN

_LRELU = 0.2


def _plan(module):
    """[(cin, hid, cout)] of the ResnetBlocks, or None when a width is outside what the tensor-core kernels take."""
    blocks = [m for m in module.resnet if hasattr(m, "conv_0")]
    plan = [(b.fin, b.fhidden, b.fout) for b in blocks]
    ok = all(c % 64 == 0 and 64 <= c <= 256 for p in plan for c in p) and module.nf == 64 and all(b.is_bias for b in blocks)
    return plan if ok else None


def use_native(module, x):
    if not RN.use_native(x):
        return False
    if _plan(module) is None:
        raise NotImplementedError("CUB ResNet: the native path covers nfilter = 64 with block widths of 64 / 128 / 192 / 256 channels "
                                  "(the reference defaults); run other widths with compute_dtype = float32 (library layers)")
    return True


def _lrelu(x, P):
    """a = actvn(x) with P rows (rows beyond x's own are zero: side inputs of the kernels are addressed up to P)."""
    a = torch.empty(P, x.shape[1], device=x.device, dtype=torch.bfloat16)
    rows = min(P, x.shape[0])
    C.check(C.lib().mv_lrelu_fwd(x.data_ptr(), a.data_ptr(), rows, x.shape[1], _LRELU, C.stream()), "mv_lrelu_fwd")
    if rows < P:
        a[rows:].zero_()
    return a


def _block_fwd(x, g, blk, tag):
    """-> (out, a, h)"""
    taps = g.taps3x3()
    a = _lrelu(x, g.P)
    xs = x if blk.wsc is None else HL.tapgemm(x, blk.wsc, 1, [0], blk.cout, g.P, geom=g, tag=f"{tag}.sc")
    h = HL.tapgemm(a, blk.w0, 9, taps, blk.hid, g.P, bias=blk.b0, act="lrelu", geom=g, tag=f"{tag}.c0")
    out = HL.tapgemm(h, blk.w1, 9, taps, blk.cout, g.P, bias=blk.b1, alpha=0.1, res=xs, geom=g, tag=f"{tag}.c1")
    return out, a, h


def _block_bwd(g_out, x, a, h, g, blk, tag, arena, need_gx=True):
    """-> (g_x, dW0, db0, dW1, db1, dWsc) with dW1 / db1 still to be scaled by 0.1 (the weight-gradient kernel reads g_out itself)."""
    taps = g.taps3x3()
    z = arena.take
    dW1, db1 = HL.wgrad(h, g_out, 9, taps, g.P, tag=f"{tag}.c1", want_db=True, dW=z(9, blk.cout, blk.hid), db=z(blk.cout))
    g_h = HL.tapgemm(g_out, blk.w1d, 9, taps, blk.hid, g.P, alpha=0.1, dact1=h, slope1=_LRELU, geom=g, tag=f"{tag}.c1d")
    dW0, db0 = HL.wgrad(a, g_h, 9, taps, g.P, tag=f"{tag}.c0", want_db=True, dW=z(9, blk.hid, blk.cin), db=z(blk.hid))
    dWsc, g_short = None, g_out
    if blk.wsc is not None:
        dWsc = HL.wgrad(x, g_out, 1, [0], g.P, tag=f"{tag}.sc", dW=z(1, blk.cout, blk.cin))
        g_short = HL.tapgemm(g_out, blk.wscd, 1, [0], blk.cin, g.P, geom=g, tag=f"{tag}.scd") if need_gx else None
    g_x = None
    if need_gx:
        g_x = HL.tapgemm(g_h, blk.w0d, 9, taps, blk.cin, g.P, dact1=a, slope1=_LRELU, res=g_short, geom=g, tag=f"{tag}.c0d")
    return g_x, dW0, db0, dW1, db1, dWsc


def _arena_floats(blocks):
    return sum(RN._block_wgrad_floats(b) for b in blocks)


def _block_params(module):
    """Flat parameter tuple of the ResnetBlocks in forward order: (w0, b0, w1, b1[, wsc]) per block, and the has-shortcut flags."""
    params, has_sc = [], []
    for b in (m for m in module.resnet if hasattr(m, "conv_0")):
        params += [b.conv_0.weight, b.conv_0.bias, b.conv_1.weight, b.conv_1.bias]
        has_sc.append(b.learned_shortcut)
        if b.learned_shortcut:
            params.append(b.conv_s.weight)
    return params, tuple(has_sc)


def _split_blocks(params, has_sc):
    out, i = [], 0
    for sc in has_sc:
        n = 5 if sc else 4
        w0, b0, w1, b1 = params[i:i + 4]
        out.append((w0, b0, w1, b1, params[i + 4] if sc else None))
        i += n
    return out, params[i:]


def _block_grads(res, scale1=0.1):
    """kernel-layout gradients of one block -> the parameters' own layouts, in parameter order."""
    g_x, dW0, db0, dW1, db1, dWsc = res
    u = HL.unpack_conv_wgrad
    out = [u(dW0, 3, 3), db0, u(dW1, 3, 3) * scale1, db1 * scale1]
    if dWsc is not None:
        out.append(u(dWsc, 1, 1))
    return out


class DecoderStackFn(torch.autograd.Function):
    """h0 (halo matrix at s0 x s0, nf0 channels, bf16) + the conv parameters -> reconstruction [n_img, 3, 64, 64] bf16
    (cub.py:226-246 after the fc layer)."""

    @staticmethod
    def forward(ctx, h0, n_img, s0, has_sc, *params):
        blocks, (wh, bh) = _split_blocks(params, has_sc)
        dev = h0.device
        B, ((whf, whd),) = RN._pack_network(blocks, [(wh, 16, wh.shape[1], True)])   # image head: 3 -> 16 output channels (zeros)
        bhp = torch.zeros(16, device=dev, dtype=torch.float32)
        bhp[: wh.shape[0]] = bh.detach().float()
        g = HL.Geom(n_img, s0, s0)
        x, saved, geoms = h0, [], []
        for i, blk in enumerate(B):
            out, a, h = _block_fwd(x, g, blk, f"cd{i}")
            saved += [x if blk.wsc is not None else a, a, h]     # the raw input is only needed for the shortcut's weight gradient
            geoms.append(g)
            if i + 1 < len(B):
                x, g = RN._upsample_fwd(out, g, blk.cout)
            else:
                x = out
        a_last = _lrelu(x, g.P)
        n_ch = wh.shape[0]
        recon = torch.empty(n_img, n_ch, g.H, g.W, device=dev, dtype=torch.bfloat16)
        HL.tapgemm(a_last, whf, 9, g.taps3x3(), 16, g.P, bias=bhp, geom=g, nchw_out=recon, n_valid=n_ch, tag="cd.head")
        ctx.save_for_backward(a_last, *saved)
        ctx.packs, ctx.whd, ctx.geoms = B, whd, geoms
        ctx.n_img, ctx.n_ch, ctx.h0_rows = n_img, n_ch, h0.shape[0]
        return recon

    @staticmethod
    def backward(ctx, g_recon):
        a_last, *saved = ctx.saved_tensors
        B, geoms = ctx.packs, ctx.geoms
        g = geoms[-1]
        dev = a_last.device
        gh = RN._pack_image(g_recon.to(torch.bfloat16).contiguous(), g)           # no activation after conv_img: plain pack
        arena = HL.ZeroArena(_arena_floats(B) + 9 * 16 * B[-1].cout + 32, dev)
        dWh, dbh = HL.wgrad(a_last, gh, 9, g.taps3x3(), g.P, tag="cd.head", want_db=True, dW=arena.take(9, 16, B[-1].cout), db=arena.take(16))
        g_out = HL.tapgemm(gh, ctx.whd, 9, g.taps3x3(), B[-1].cout, g.P, dact1=a_last, slope1=_LRELU, geom=g, tag="cd.head.d")
        grads = [None] * len(B)
        g_h0 = None
        for i in range(len(B) - 1, -1, -1):
            x, a, h = saved[3 * i:3 * i + 3]
            need = i > 0 or ctx.needs_input_grad[0]
            res = _block_bwd(g_out, x, a, h, geoms[i], B[i], f"cd{i}", arena, need_gx=need)
            grads[i] = _block_grads(res)
            if i > 0:
                gc = geoms[i - 1]
                g_out = torch.empty(gc.P, B[i].cin, device=dev, dtype=torch.bfloat16)
                C.check(C.lib().mv_upsample2x_bwd(res[0].data_ptr(), None, g_out.data_ptr(), None, gc.n_img, gc.H, gc.W, B[i].cin, 1.0,
                                                  _LRELU, C.stream()), "mv_upsample2x_bwd")
            else:
                g_h0 = res[0]
        if g_h0 is not None:
            g_h0 = g_h0[: ctx.h0_rows]
        flat = [t for blk in grads for t in blk]
        flat += [HL.unpack_conv_wgrad(dWh, 3, 3)[: ctx.n_ch], dbh[: ctx.n_ch]]
        return (g_h0, None, None, None) + tuple(flat)


def decoder_forward(dec, z):
    zz = z.reshape(-1, z.size(-1))
    n_img = zz.shape[0]
    s0 = dec.s0
    S = (s0 + 1) * (s0 + 1)
    h0 = RN.FcHaloFn.apply(zz, dec.fc.weight, dec.fc.bias, s0, dec.nf0).view(n_img * S, dec.nf0)
    params, has_sc = _block_params(dec)
    recon = DecoderStackFn.apply(h0, n_img, s0, has_sc, *params, dec.conv_img.weight, dec.conv_img.bias)
    return ModelOutput(reconstruction=recon.view(*z.size()[:-1], *recon.shape[1:]))


class EncoderStackFn(torch.autograd.Function):
    """x [n, 3, 64, 64] -> actvn(features) as a halo matrix [n * (s0+1)^2, nf0] (cub.py:185-189 before the fc heads)."""

    @staticmethod
    def forward(ctx, x, has_sc, wi, bi, *params):
        blocks, _rest = _split_blocks(params, has_sc)
        n_img, size = x.shape[0], x.shape[2]
        g = HL.Geom(n_img, size, size)
        x16 = RN._pack_image(x.detach().to(torch.bfloat16).contiguous(), g)
        B, ((wif, _),) = RN._pack_network(blocks, [(wi, wi.shape[0], 16, False)])     # image conv: 3 -> 16 input channels (zeros)
        cur = HL.tapgemm(x16, wif, 9, g.taps3x3(), wi.shape[0], g.P, bias=bi.detach().float().contiguous(), geom=g, tag="ce.img")
        saved, geoms = [], []
        for i, blk in enumerate(B):
            out, a, h = _block_fwd(cur, g, blk, f"ce{i}")
            saved += [cur if blk.wsc is not None else a, a, h]
            geoms.append(g)
            if i + 1 < len(B):
                cur, g = RN._avgpool_fwd(out, g, blk.cout)
            else:
                cur = out
        af = _lrelu(cur, g.P)
        ctx.save_for_backward(x16, cur, *saved)   # `cur` (the raw features) carries the sign that lrelu' needs
        ctx.packs, ctx.geoms = B, geoms
        ctx.n_img, ctx.cin_img = n_img, wi.shape[1]
        return af[: n_img * g.S]

    @staticmethod
    def backward(ctx, g_af):
        x16, feat, *saved = ctx.saved_tensors
        B, geoms = ctx.packs, ctx.geoms
        g = geoms[-1]
        dev = x16.device
        gp = torch.zeros(g.P, B[-1].cout, device=dev, dtype=torch.bfloat16)
        gp[: g_af.shape[0]].copy_(g_af)
        g_out = torch.empty_like(gp)
        C.check(C.lib().mv_scale_dact(gp.data_ptr(), feat.data_ptr(), g_out.data_ptr(), g.P, B[-1].cout, 1.0, _LRELU, C.stream()), "mv_scale_dact")
        arena = HL.ZeroArena(_arena_floats(B), dev)
        grads = [None] * len(B)
        for i in range(len(B) - 1, -1, -1):
            x, a, h = saved[3 * i:3 * i + 3]
            res = _block_bwd(g_out, x, a, h, geoms[i], B[i], f"ce{i}", arena)
            grads[i] = _block_grads(res)
            if i > 0:
                gf = geoms[i - 1]       # the finer stage the pooling read from
                g_out = torch.empty(gf.P, B[i].cin, device=dev, dtype=torch.bfloat16)
                C.check(C.lib().mv_avgpool3s2_bwd(res[0].data_ptr(), None, g_out.data_ptr(), None, gf.n_img, gf.H, gf.W, B[i].cin, 1.0,
                                                  _LRELU, C.stream()), "mv_avgpool3s2_bwd")
            else:
                g_a0 = res[0]
        g0 = geoms[0]
        # conv_img (3 -> 64): weight gradient with the operand roles swapped (the 16-channel image is the N side)
        dWs = HL.wgrad(g_a0, x16, 9, [-o for o in g0.taps3x3()], g0.P, tag="ce.img")       # [9, 16, 64]
        dWi = dWs.view(3, 3, 16, dWs.shape[2]).permute(3, 2, 0, 1)[:, : ctx.cin_img]
        dbi = RN._colsum(g_a0, g0.P, dWs.shape[2])
        return (None, None, dWi, dbi) + tuple(t for blk in grads for t in blk)


def encoder_forward(enc, x):
    params, has_sc = _block_params(enc)
    h = EncoderStackFn.apply(x, has_sc, enc.conv_img.weight, enc.conv_img.bias, *params)
    mu, lv = RN.FcFromHaloFn.apply(h, enc.fc_mu.weight, enc.fc_mu.bias, enc.fc_logvar.weight, enc.fc_logvar.bias, enc.s0)
    return ModelOutput(embedding=mu, log_covariance=lv)
