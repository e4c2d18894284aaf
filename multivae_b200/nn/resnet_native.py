"""Native (tcgen05) execution of the PolyMNIST ResNet decoder (reference: models/nn/mmnist.py:214-251,321-366).

`DecoderResnetMMNIST.forward` routes here on CUDA when the model computes in bf16.  The whole convolutional
stack (3 ResnetBlocks, 2 nearest upsamplings, the image head) runs as hand-written sm_100a kernels through
the C-ABI — mv_tapgemm (forward + data gradients), mv_wgrad (weight gradients) and the HBM-bound helpers
of csrc/halo_ops.cu — on activations kept in the shared-halo NHWC bf16 layout; fp32 master weights are
packed to bf16 tap-major matrices once per step.  The fully connected input layer (1.6 of 213 MFLOP per
image) runs on the general tcgen05 GEMM (mv_gemm); its permuted weight makes it emit the halo layout directly.

There is no fallback inside this path: if the C-ABI library is missing, it raises.
"""
import os

import torch
import torch.nn.functional as F

from .. import _cabi as C
from ..containers import ModelOutput
from . import functional as NF
from . import halo as HL

_LRELU = 0.2
_MASK_D = os.environ.get("MULTIVAE_B200_MASK_D", "1") != "0"   # save the last block's `d` as a sign mask
_SPLIT_C0D = os.environ.get("MULTIVAE_B200_SPLIT_C0D", "1") != "0"
_FUSE_SCD = os.environ.get("MULTIVAE_B200_FUSE_SCD", "1") != "0"   # shortcut data gradient fused into the conv0 data-gradient launches
_MASK_H = os.environ.get("MULTIVAE_B200_MASK_H", "1") != "0"   # sign mask of `h` next to h itself (64-channel hidden layers)


def use_native(x):
    """Native path: CUDA tensor, backend not forced to torch, bf16 compute (autocast active)."""
    if not x.is_cuda or NF.backend() == "torch":
        return False
    return NF.backend() == "native" or torch.is_autocast_enabled()


def status(which):
    return "native sm_100a: tcgen05 tap-GEMM conv fwd/dgrad + tcgen05 wgrad + halo kernels; fc layers on the tcgen05 GEMM (mv_gemm)"


# ---- low-level wrappers over the helper kernels -------------------------------------------------
def _upsample_fwd(x, g, c):
    go = g.up2()
    out = torch.empty(go.P, c, device=x.device, dtype=torch.bfloat16)
    C.check(C.lib().mv_upsample2x_fwd(x.data_ptr(), out.data_ptr(), g.n_img, g.H, g.W, c, C.stream()), "mv_upsample2x_fwd")
    return out, go


def _upsample_bwd(g_out, act, g, c, alpha):
    """g: coarse geometry.  Returns (g_in, g_pre)."""
    g_in = torch.empty(g.P, c, device=g_out.device, dtype=torch.bfloat16)
    g_pre = torch.empty_like(g_in)
    C.check(C.lib().mv_upsample2x_bwd(g_out.data_ptr(), act.data_ptr(), g_in.data_ptr(), g_pre.data_ptr(), g.n_img, g.H, g.W, c,
                                      float(alpha), _LRELU, C.stream()), "mv_upsample2x_bwd")
    return g_in, g_pre


def _colsum(G, P, n, out=None):
    if out is None:
        out = torch.zeros(n, device=G.device, dtype=torch.float32)
    C.check(C.lib().mv_colsum(G.data_ptr(), min(P, G.shape[0]), G.stride(0), n, out.data_ptr(), C.stream()), "mv_colsum")
    return out


def _avgpool_fwd(x, g, c):
    go = HL.Geom(g.n_img, g.H // 2, g.W // 2)
    out = torch.empty(go.P, c, device=x.device, dtype=torch.bfloat16)
    C.check(C.lib().mv_avgpool3s2_fwd(x.data_ptr(), out.data_ptr(), g.n_img, g.H, g.W, c, C.stream()), "mv_avgpool3s2_fwd")
    return out, go


def _avgpool_bwd(g_out, act, g, c, alpha):
    """g: fine (input) geometry.  Returns (g_in, g_pre = alpha * g_in * lrelu'(act))."""
    g_in = torch.empty(g.P, c, device=g_out.device, dtype=torch.bfloat16)
    g_pre = torch.empty_like(g_in)
    C.check(C.lib().mv_avgpool3s2_bwd(g_out.data_ptr(), act.data_ptr(), g_in.data_ptr(), g_pre.data_ptr(), g.n_img, g.H, g.W, c,
                                      float(alpha), _LRELU, C.stream()), "mv_avgpool3s2_bwd")
    return g_in, g_pre


def _pack_image(x_bf16, g, y_out=None):
    """dense NCHW bf16 [n, ch<=16, H, W] (optionally times lrelu'(y_out)) -> halo matrix [P, 16]."""
    out = torch.empty(g.P, 16, device=x_bf16.device, dtype=torch.bfloat16)
    C.check(C.lib().mv_head_grad_pack(x_bf16.data_ptr(), None if y_out is None else y_out.data_ptr(), out.data_ptr(), g.n_img, g.H,
                                      g.W, x_bf16.shape[1], _LRELU, C.stream()), "mv_head_grad_pack")
    return out


class _Block:
    """bf16 packs of one ResnetBlock's weights (forward and data-gradient orientation)."""

    def __init__(self, w0, b0, w1, b1, wsc, packs):
        self.cin, self.hid, self.cout = w0.shape[1], w0.shape[0], w1.shape[0]
        (self.w0, self.w0d), (self.w1, self.w1d) = packs[0], packs[1]
        self.b0, self.b1 = b0.detach().float().contiguous(), b1.detach().float().contiguous()
        self.wsc = self.wscd = None
        if wsc is not None:
            self.wsc, self.wscd = packs[2]
        self.w0d_halves = None   # data-gradient packs of conv0 per 64 input channels (two 64-output launches, see _block_bwd)


def _pack_network(blocks, extra):
    """One mv_pack_conv_weights launch for a whole stack: `blocks` = [(w0, b0, w1, b1, wsc | None)], `extra` = further
    (weight, Npad, Cpad, want_dgrad) specs.  Returns ([_Block], [packs of the extra specs])."""
    specs = []
    for w0, b0, w1, b1, wsc in blocks:
        specs += [(w0, w0.shape[0], w0.shape[1], True), (w1, w1.shape[0], w1.shape[1], True)]
        if wsc is not None:
            specs.append((wsc, wsc.shape[0], wsc.shape[1], True))
    # conv0 of a 128 -> 64 block: its data gradient (64 -> 128) runs as two 64-output convolutions on the three-taps-per-MMA
    # kernel (the 128-column tap-GEMM streams its 144 KB of weights per tile), each with its own pack
    split = [k for k, (w0, *_r) in enumerate(blocks) if _SPLIT_C0D and w0.shape[0] == 64 and w0.shape[1] == 128]
    halves = []
    for k in split:
        w0 = blocks[k][0].detach()
        halves += [(w0[:, :64].contiguous(), 64, 64, True), (w0[:, 64:].contiguous(), 64, 64, True)]
    n_extra = len(extra)
    packs = HL.pack_conv_weights(specs + list(extra) + halves)
    out, i = [], 0
    for w0, b0, w1, b1, wsc in blocks:
        n = 2 if wsc is None else 3
        out.append(_Block(w0, b0, w1, b1, wsc, packs[i:i + n]))
        i += n
    for j, k in enumerate(split):
        out[k].w0d_halves = (packs[i + n_extra + 2 * j][1], packs[i + n_extra + 2 * j + 1][1])
    return out, packs[i:i + n_extra]


def _block_fwd(x, g, blk, tag, mask_d=False):
    """x_s + 0.1 * lrelu(conv1(lrelu(conv0(x))))  ->  (out, saved h, saved d).
    mask_d: `d` (needed only for the sign of the LeakyReLU derivative) is saved as one bit per element (int64 word per row)."""
    taps = g.taps3x3()
    xs = x if blk.wsc is None else HL.tapgemm(x, blk.wsc, 1, [0], blk.cout, g.P, geom=g, tag=f"{tag}.sc")
    hm = None
    if _MASK_H and blk.hid == 64 and blk.cin in (64, 128) and blk.cout in (64, 128):
        # sign bits of h for the LeakyReLU derivative in the data gradient of conv1 (instead of re-reading h as a side tile)
        hm = torch.empty(HL.mask_rows(g.P), device=x.device, dtype=torch.int64)
    h = HL.tapgemm(x, blk.w0, 9, taps, blk.hid, g.P, bias=blk.b0, act="lrelu", geom=g, tag=f"{tag}.c0", out2_mask=hm)
    if mask_d:
        assert blk.cout == 64
        d = torch.empty(HL.mask_rows(g.P), device=x.device, dtype=torch.int64)
        out = HL.tapgemm(h, blk.w1, 9, taps, blk.cout, g.P, bias=blk.b1, act="lrelu", alpha=0.1, res=xs, out2_mask=d, geom=g,
                         tag=f"{tag}.c1")
        return out, h, d, hm
    d = torch.empty(g.P, blk.cout, device=x.device, dtype=torch.bfloat16)
    out = HL.tapgemm(h, blk.w1, 9, taps, blk.cout, g.P, bias=blk.b1, act="lrelu", alpha=0.1, res=xs, out2=d, out2_pre=True,
                     geom=g, tag=f"{tag}.c1")
    return out, h, d, hm


def _block_wgrad_floats(blk):
    """fp32 elements of a block's weight/bias gradients (ZeroArena sizing; +4 per tensor for alignment)."""
    n = 9 * blk.cout * blk.hid + blk.cout + 9 * blk.hid * blk.cin + blk.hid + 16
    if blk.wsc is not None:
        n += blk.cout * blk.cin + 4
    return n


def _grad_mode():
    """How parameter gradients leave the native backward passes (MULTIVAE_B200_DIRECT_GRADS):
    "1" (default)  one mv_unpack_wgrad_add launch per stack adds the [T][N][C] buffers into the parameters' .grad tensors
                   (when those exist as contiguous fp32 tensors — the trainer's flat buffer), None goes back to autograd;
    "nct"          the weight-gradient kernels accumulate straight into .grad with strided atomics (mv_wgrad_nct);
    "0"            gradients are returned to autograd (one permuted `grad += dW` kernel per parameter)."""
    return {"0": "autograd", "nct": "nct"}.get(os.environ.get("MULTIVAE_B200_DIRECT_GRADS", "1"), "unpack")


_DIRECT = [False]


class direct_grads:
    """Context manager the trainer wraps around `loss.backward()`: inside it the native stacks may add their weight gradients
    straight into the parameters' `.grad` tensors (the trainer's flat all-reduce buffer) and hand `None` back to autograd.
    Outside it (plain `loss.backward()`, `torch.autograd.grad`, gradient hooks, `backward(inputs=...)`) every gradient is
    returned through autograd like any other Function."""

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.prev, _DIRECT[0] = _DIRECT[0], self.on

    def __exit__(self, *a):
        _DIRECT[0] = self.prev


def _direct_targets(params):
    """The parameters' own .grad tensors if the trainer opted in (`direct_grads`) and ALL of them can be written in place
    (require a gradient; fp32, contiguous, allocated), else None."""
    if not _DIRECT[0] or _grad_mode() == "autograd" or torch.is_grad_enabled():
        return None
    tg = []
    for p in params:
        g = getattr(p, "grad", None)
        if not p.requires_grad:
            return None
        if g is None or g.dtype != torch.float32 or not g.is_cuda or not g.is_contiguous() or g.shape != p.shape:
            return None
        tg.append(g)
    return tg


def _block_bwd(g_out, g_dpre, x, h, g, blk, tag, need_gx=True, arena=None, tg=None, hm=None, g_out_from_dpre=None):
    """g_out: gradient of the block output; g_dpre = 0.1 * g_out * lrelu'(d) (produced upstream).
    g_out_from_dpre = sign mask of d: g_out is NOT materialised (None); the skip connection recovers it in the conv0 data-gradient
    epilogue as g_dpre * (bit ? 10 : 50) (blocks without a shortcut convolution, 64 channels).
    Returns (g_x, dW0, db0, dW1, db1, dWsc); with tg = (w0.grad, b0.grad, w1.grad, b1.grad, wsc.grad | None) the weight / bias
    gradients are accumulated into those tensors by the kernels and None is returned in their place."""
    taps = g.taps3x3()
    z = (lambda *s: None) if arena is None else arena.take
    if tg is not None:
        HL.wgrad(h, g_dpre, 9, taps, g.P, tag=f"{tag}.c1", want_db=True, grad_out=tg[2], db=tg[3])
        dW1 = db1 = None
    else:
        dW1, db1 = HL.wgrad(h, g_dpre, 9, taps, g.P, tag=f"{tag}.c1", want_db=True, dW=z(9, blk.cout, blk.hid), db=z(blk.cout))
    dkw = dict(dact1=h) if hm is None else dict(dmask1=hm)
    g_hpre = HL.tapgemm(g_dpre, blk.w1d, 9, taps, blk.hid, g.P, slope1=_LRELU, geom=g, tag=f"{tag}.c1d", **dkw)
    if tg is not None:
        HL.wgrad(x, g_hpre, 9, taps, g.P, tag=f"{tag}.c0", want_db=True, grad_out=tg[0], db=tg[1])
        dW0 = db0 = None
    else:
        dW0, db0 = HL.wgrad(x, g_hpre, 9, taps, g.P, tag=f"{tag}.c0", want_db=True, dW=z(9, blk.hid, blk.cin), db=z(blk.hid))
    dWsc = None
    g_short = g_out
    # 128 -> 64 blocks: the shortcut's data gradient rides on the conv0 data-gradient launches as a fused 1x1 term (no g_short tensor)
    fuse_sc = need_gx and blk.w0d_halves is not None and blk.wsc is not None and blk.cout == 64 and _FUSE_SCD
    if blk.wsc is not None:
        if tg is not None:
            HL.wgrad(x, g_out, 1, [0], g.P, tag=f"{tag}.sc", grad_out=tg[4])
        else:
            dWsc = HL.wgrad(x, g_out, 1, [0], g.P, tag=f"{tag}.sc", dW=z(1, blk.cout, blk.cin))
        if not fuse_sc:
            g_short = HL.tapgemm(g_out, blk.wscd, 1, [0], blk.cin, g.P, geom=g, tag=f"{tag}.scd") if need_gx else None
    g_x = None
    if fuse_sc:
        g_x = torch.empty(g.P, blk.cin, device=g_hpre.device, dtype=torch.bfloat16)
        for j, wd in enumerate(blk.w0d_halves):
            HL.tapgemm(g_hpre, wd, 9, taps, 64, g.P, a2=g_out, w2=blk.wscd[64 * j:64 * j + 64], out=g_x[:, 64 * j:64 * j + 64], geom=g,
                       tag=f"{tag}.c0d")
    elif need_gx and blk.w0d_halves is not None:
        g_x = torch.empty(g.P, blk.cin, device=g_hpre.device, dtype=torch.bfloat16)
        for j, wd in enumerate(blk.w0d_halves):
            HL.tapgemm(g_hpre, wd, 9, taps, 64, g.P, res=g_short[:, 64 * j:64 * j + 64], out=g_x[:, 64 * j:64 * j + 64], geom=g,
                       tag=f"{tag}.c0d")
    elif need_gx and g_out_from_dpre is not None:
        assert blk.wsc is None
        g_x = HL.tapgemm(g_hpre, blk.w0d, 9, taps, blk.cin, g.P, res=g_dpre, res_mask=g_out_from_dpre,
                         res_scale=(1.0 / 0.1, 1.0 / (0.1 * _LRELU)), geom=g, tag=f"{tag}.c0d")
    elif need_gx:
        g_x = HL.tapgemm(g_hpre, blk.w0d, 9, taps, blk.cin, g.P, res=g_short, geom=g, tag=f"{tag}.c0d")
    return g_x, dW0, db0, dW1, db1, dWsc


class DecoderStackFn(torch.autograd.Function):
    """h0 (halo matrix at 7x7, 256 ch, bf16) + the 17 conv parameters -> reconstruction [n_img, 3, 28, 28] bf16."""

    @staticmethod
    def forward(ctx, h0, n_img, *params):
        (w10, b10, w11, b11, wsc1, w20, b20, w21, b21, wsc2, w30, b30, w31, b31, wh, bh) = params
        dev = h0.device
        (B1, B2, B3), ((whf, whd),) = _pack_network([(w10, b10, w11, b11, wsc1), (w20, b20, w21, b21, wsc2), (w30, b30, w31, b31, None)],
                                                    [(wh, 16, wh.shape[1], True)])   # image head: 3 -> 16 output channels (zeros)
        bhp = torch.zeros(16, device=dev, dtype=torch.float32)
        bhp[: wh.shape[0]] = bh.detach().float()
        g7 = HL.Geom(n_img, 7, 7)
        o1, h1, d1, _ = _block_fwd(h0, g7, B1, "b1")
        u1, g14 = _upsample_fwd(o1, g7, B1.cout)
        o2, h2, d2, hm2 = _block_fwd(u1, g14, B2, "b2")
        u2, g28 = _upsample_fwd(o2, g14, B2.cout)
        o3, h3, d3, hm3 = _block_fwd(u2, g28, B3, "b3", mask_d=_MASK_D)   # d3: sign bits only (1/16 of the bf16 tensor's HBM traffic)
        n_ch = wh.shape[0]
        recon = torch.empty(n_img, n_ch, 28, 28, device=dev, dtype=torch.bfloat16)
        HL.tapgemm(o3, whf, 9, g28.taps3x3(), 16, g28.P, bias=bhp, act="lrelu", geom=g28, nchw_out=recon, n_valid=n_ch, tag="head")
        ctx.save_for_backward(h0, h1, d1, u1, h2, d2, u2, h3, d3, o3, recon, hm2, hm3)
        ctx.packs = (B1, B2, B3, whd)
        ctx.params = params
        ctx.n_img, ctx.n_ch = n_img, n_ch
        return recon

    @staticmethod
    def backward(ctx, g_recon):
        h0, h1, d1, u1, h2, d2, u2, h3, d3, o3, recon, hm2, hm3 = ctx.saved_tensors
        B1, B2, B3, whd = ctx.packs
        n_img, n_ch = ctx.n_img, ctx.n_ch
        lib = C.lib()
        g7 = HL.Geom(n_img, 7, 7)
        g14, g28 = g7.up2(), g7.up2().up2()
        g_recon = g_recon.to(torch.bfloat16).contiguous()
        # image head: gradient through the LeakyReLU, then weight / bias / data gradients
        gh = torch.empty(g28.P, 16, device=h0.device, dtype=torch.bfloat16)
        C.check(lib.mv_head_grad_pack(g_recon.data_ptr(), recon.data_ptr(), gh.data_ptr(), n_img, 28, 28, n_ch, _LRELU, C.stream()),
                "mv_head_grad_pack")
        tgt = _direct_targets(ctx.params)  # (w10, b10, w11, b11, wsc1, w20, b20, w21, b21, wsc2, w30, b30, w31, b31, wh, bh).grad
        tg = tgt if _grad_mode() == "nct" else None
        arena = None
        if tg is not None:
            HL.wgrad(o3, gh, 9, g28.taps3x3(), g28.P, tag="head", want_db=True, grad_out=tg[14], n_valid=n_ch, db=tg[15])
            dWh = dbh = None
        else:
            arena = HL.ZeroArena(sum(_block_wgrad_floats(b) for b in (B1, B2, B3)) + 9 * 16 * B3.cout + 32, h0.device)
            dWh, dbh = HL.wgrad(o3, gh, 9, g28.taps3x3(), g28.P, tag="head", want_db=True, dW=arena.take(9, 16, B3.cout), db=arena.take(16))
        T = (lambda i, j: None) if tg is None else (lambda i, j: tuple(tg[i:j]) + ((None,) if j - i == 4 else ()))
        if d3.dtype == torch.int64 and B3.w0d_halves is None:
            # ONE output: the pre-scaled gradient 0.1 * g * lrelu'(d3); the block's skip connection un-scales it from d3's sign
            # bits in the conv0 data-gradient epilogue (saves writing and re-reading a 64-channel 28x28 tensor)
            g_dpre3 = HL.tapgemm(gh, whd, 9, g28.taps3x3(), B3.cout, g28.P, alpha=0.1, dmask1=d3, slope1=_LRELU, geom=g28, tag="head.d")
            g_u2, dW30, db30, dW31, db31, _ = _block_bwd(None, g_dpre3, u2, h3, g28, B3, "b3", arena=arena, tg=T(10, 14), hm=hm3,
                                                         g_out_from_dpre=d3)
        else:
            g_dpre3 = torch.empty(g28.P, B3.cout, device=h0.device, dtype=torch.bfloat16)
            dkw = dict(dmask2=d3) if d3.dtype == torch.int64 else dict(dact2=d3)
            g_o3 = HL.tapgemm(gh, whd, 9, g28.taps3x3(), B3.cout, g28.P, out2=g_dpre3, alpha2=0.1, slope2=_LRELU, geom=g28, tag="head.d",
                              **dkw)
            g_u2, dW30, db30, dW31, db31, _ = _block_bwd(g_o3, g_dpre3, u2, h3, g28, B3, "b3", arena=arena, tg=T(10, 14), hm=hm3)
        g_o2, g_dpre2 = _upsample_bwd(g_u2, d2, g14, B2.cout, 0.1)
        g_u1, dW20, db20, dW21, db21, dWsc2 = _block_bwd(g_o2, g_dpre2, u1, h2, g14, B2, "b2", arena=arena, tg=T(5, 10), hm=hm2)
        g_o1, g_dpre1 = _upsample_bwd(g_u1, d1, g7, B1.cout, 0.1)
        need_h0 = ctx.needs_input_grad[0]
        g_h0, dW10, db10, dW11, db11, dWsc1 = _block_bwd(g_o1, g_dpre1, h0, h1, g7, B1, "b1", need_gx=need_h0, arena=arena, tg=T(0, 5))
        if g_h0 is not None:
            g_h0 = g_h0[: h0.shape[0]]
        if tg is not None:   # every parameter gradient was accumulated in place by the kernels
            return (g_h0, None) + (None,) * 16
        if tgt is not None:  # one launch adds all sixteen buffers into the .grad tensors
            HL.unpack_wgrad_add(list(zip((dW10, db10, dW11, db11, dWsc1, dW20, db20, dW21, db21, dWsc2, dW30, db30, dW31, db31, dWh, dbh),
                                         tgt)))
            return (g_h0, None) + (None,) * 16
        u = HL.unpack_conv_wgrad
        grads = (u(dW10, 3, 3), db10, u(dW11, 3, 3), db11, u(dWsc1, 1, 1), u(dW20, 3, 3), db20, u(dW21, 3, 3), db21,
                 u(dWsc2, 1, 1), u(dW30, 3, 3), db30, u(dW31, 3, 3), db31, u(dWh, 3, 3)[:n_ch], dbh[:n_ch])
        return (g_h0, None) + grads


_PERM_CACHE = {}


def _fc_perm(device, s0=7, ch=256):
    """(gather, scatter) for an s0 x s0 x ch feature map: `gather` [(s0+1)^2 * ch] maps halo feature (y'*(s0+1) + x)*ch + c to fc
    row c*s0^2 + (y'-1)*s0 + x (halo positions -> the extra all-zero row ch*s0^2); `scatter` [ch*s0^2] is its inverse on the
    real rows."""
    key = (str(device), s0, ch)
    if key not in _PERM_CACHE:
        n_real = ch * s0 * s0
        idx = torch.full((s0 + 1, s0 + 1, ch), n_real, dtype=torch.long)
        c = torch.arange(ch)
        yy = torch.arange(s0).view(s0, 1, 1)
        xx = torch.arange(s0).view(1, s0, 1)
        idx[1:, :s0] = c.view(1, 1, ch) * (s0 * s0) + yy * s0 + xx
        idx = idx.reshape(-1)
        inv = torch.empty(n_real, dtype=torch.long)
        pos = torch.nonzero(idx < n_real).squeeze(1)
        inv[idx[pos]] = pos
        _PERM_CACHE[key] = (idx.to(device), inv.to(device))
    return _PERM_CACHE[key]


def _gather_cast(src, idx, mode, out_rows, out_cols):
    """fp32 matrix -> bf16 GEMM operand [out_rows, out_cols] gathered by `idx` along rows (mode 0) or columns (mode 1), with a
    16-byte row pitch (mv_gather_cast: one launch instead of cat + cast + index kernels)."""
    src = src.detach()
    assert src.dtype == torch.float32 and src.stride(1) == 1 and idx.dtype == torch.int64
    ld = (out_cols + 7) // 8 * 8
    out = torch.empty(out_rows, ld, device=src.device, dtype=torch.bfloat16)
    C.check(C.lib().mv_gather_cast(src.data_ptr(), src.shape[0], src.shape[1], src.stride(0), idx.data_ptr(), out_rows, out_cols, mode,
                                   out.data_ptr(), ld, C.stream()), "mv_gather_cast")
    return out[:, :out_cols]


def _gather_f32(src, src_cols, idx, mode, out_rows, out_cols, into=None):
    """fp32 gather (see _gather_cast) of the first src_cols columns of `src`; `into`: accumulate into that contiguous tensor
    (a parameter's .grad) instead of returning a new one."""
    src = src.detach()
    if src.dim() == 1:
        src = src.unsqueeze(1)
    assert src.dtype == torch.float32 and src.stride(1) == 1
    out = torch.empty(out_rows, out_cols, device=src.device, dtype=torch.float32) if into is None else into
    assert out.is_contiguous() and out.numel() == out_rows * out_cols
    C.check(C.lib().mv_gather_f32(src.data_ptr(), src.shape[0], src_cols, src.stride(0), idx.data_ptr(), out_rows, out_cols, mode,
                                  out.data_ptr(), out_cols, 0.0 if into is None else 1.0, C.stream()), "mv_gather_f32")
    return out


class FcHaloFn(torch.autograd.Function):
    """fc of the decoder emitting the 7x7 halo matrix directly: rows of the weight are gathered into halo order (zero rows at
    halo positions), so `z @ Wp^T + bp` IS the [n_img*64, 256] activation matrix.  Native GEMMs (mv_gemm: forward with the bias in
    the epilogue, data gradient with W read MN-major, weight gradient with both operands MN-major); the backward gathers the real
    rows back (no scatter-add)."""

    @staticmethod
    def forward(ctx, z, weight, bias, s0=7, ch=256):
        from .linear_native import _to_bf16_padded, gemm
        gather, scatter = _fc_perm(z.device, s0, ch)
        F_ = gather.numel()
        wp = _gather_cast(weight.float(), gather, 0, F_, weight.shape[1])          # [(s0+1)^2 * ch, L] bf16, halo order
        bp = _gather_f32(bias.float(), 1, gather, 0, F_, 1).view(F_)
        zb = _to_bf16_padded(z)
        n, L = zb.shape
        h0 = torch.empty(n, F_, device=z.device, dtype=torch.bfloat16)
        gemm(zb, wp, n, F_, L, h0, bias=bp, tag="dec.fc")
        ctx.save_for_backward(zb, wp)
        ctx.scatter = scatter
        ctx.params = (weight, bias)
        return h0

    @staticmethod
    def backward(ctx, g):
        from .linear_native import gemm
        zb, wp = ctx.saved_tensors
        n, L = zb.shape
        F_ = wp.shape[0]
        g = g.reshape(n, F_).to(torch.bfloat16).contiguous()
        gz = None
        if ctx.needs_input_grad[0]:
            gz = torch.empty(n, L, device=g.device, dtype=torch.float32)
            gemm(g, wp, n, L, F_, gz, b_mn=True, out_kind=1, tag="dec.fc.d")
        gw = torch.zeros(F_, L, device=g.device, dtype=torch.float32)
        gemm(g, zb, F_, L, n, gw, a_mn=True, b_mn=True, out_kind=2, tag="dec.fc.w")
        gb = torch.zeros(F_, device=g.device, dtype=torch.float32)
        C.check(C.lib().mv_colsum_any(g.data_ptr(), n, F_, F_, gb.data_ptr(), C.stream()), "mv_colsum_any")
        tgt = _direct_targets(ctx.params)
        R = ctx.scatter.numel()
        if tgt is not None:   # trainer opt-in: straight into the flat gradient buffer (final when this backward returns)
            _gather_f32(gw, L, ctx.scatter, 0, R, L, into=tgt[0])
            _gather_f32(gb, 1, ctx.scatter, 0, R, 1, into=tgt[1])
            return gz, None, None, None, None
        return gz, _gather_f32(gw, L, ctx.scatter, 0, R, L), _gather_f32(gb, 1, ctx.scatter, 0, R, 1).view(R), None, None


def decoder_params(dec):
    r = dec.resnet
    b1, b2, b3 = r[0], r[2], r[4]
    return (b1.conv_layers[0].weight, b1.conv_layers[0].bias, b1.conv_layers[2].weight, b1.conv_layers[2].bias,
            b1.shortcut_layer.weight,
            b2.conv_layers[0].weight, b2.conv_layers[0].bias, b2.conv_layers[2].weight, b2.conv_layers[2].bias,
            b2.shortcut_layer.weight,
            b3.conv_layers[0].weight, b3.conv_layers[0].bias, b3.conv_layers[2].weight, b3.conv_layers[2].bias,
            dec.conv_img[0].weight, dec.conv_img[0].bias)


def decoder_forward(dec, z, out_dtype=None):
    zz = z.reshape(-1, z.size(-1))
    n_img = zz.shape[0]
    h0 = FcHaloFn.apply(zz, dec.fc.weight, dec.fc.bias, 7, 256).view(n_img * 64, 256)  # halo matrix at 7x7
    recon = DecoderStackFn.apply(h0, n_img, *decoder_params(dec))
    recon = recon.view(*z.size()[:-1], *recon.shape[1:])
    return ModelOutput(reconstruction=recon)


class EncoderStackFn(torch.autograd.Function):
    """One branch (shared `u` or private `w`) of EncoderResnetMMNIST up to the flatten: x [n, 3, 28, 28] ->
    halo matrix [n*64, 256] at 7x7 (mmnist.py:254-300): conv_img, ResnetBlock(64,64), AvgPool, ResnetBlock(64,128),
    AvgPool, ResnetBlock(128,256)."""

    @staticmethod
    def forward(ctx, x, *params):
        (wi, bi, w10, b10, w11, b11, w20, b20, w21, b21, wsc2, w30, b30, w31, b31, wsc3) = params
        n_img = x.shape[0]
        g28 = HL.Geom(n_img, 28, 28)
        x16 = _pack_image(x.detach().to(torch.bfloat16).contiguous(), g28)
        (B1, B2, B3), ((wif, _),) = _pack_network([(w10, b10, w11, b11, None), (w20, b20, w21, b21, wsc2), (w30, b30, w31, b31, wsc3)],
                                                  [(wi, wi.shape[0], 16, False)])   # image conv: 3 -> 16 input channels (zeros)
        a0 = HL.tapgemm(x16, wif, 9, g28.taps3x3(), wi.shape[0], g28.P, bias=bi.detach().float().contiguous(), geom=g28, tag="e.img")
        o1, h1, d1, hm1 = _block_fwd(a0, g28, B1, "e1")
        x2, g14 = _avgpool_fwd(o1, g28, B1.cout)
        o2, h2, d2, hm2 = _block_fwd(x2, g14, B2, "e2")
        x3, g7 = _avgpool_fwd(o2, g14, B2.cout)
        o3, h3, d3, _ = _block_fwd(x3, g7, B3, "e3")
        ctx.save_for_backward(x16, a0, h1, d1, x2, h2, d2, x3, h3, d3, hm1, hm2)
        ctx.packs = (B1, B2, B3)
        ctx.params = params
        ctx.n_img, ctx.cin_img = n_img, wi.shape[1]
        return o3[: n_img * 64]

    @staticmethod
    def backward(ctx, g_o3):
        x16, a0, h1, d1, x2, h2, d2, x3, h3, d3, hm1, hm2 = ctx.saved_tensors
        B1, B2, B3 = ctx.packs
        n_img = ctx.n_img
        g28 = HL.Geom(n_img, 28, 28)
        g14, g7 = HL.Geom(n_img, 14, 14), HL.Geom(n_img, 7, 7)
        lib = C.lib()
        g3 = torch.zeros(g7.P, B3.cout, device=x16.device, dtype=torch.bfloat16)
        g3[: n_img * 64].copy_(g_o3)
        g_d3pre = torch.empty_like(g3)
        C.check(lib.mv_scale_dact(g3.data_ptr(), d3.data_ptr(), g_d3pre.data_ptr(), g7.P, B3.cout, 0.1, _LRELU, C.stream()),
                "mv_scale_dact")
        # block parameters (w10, b10, w11, b11 | w20, b20, w21, b21, wsc2 | w30, b30, w31, b31, wsc3) = ctx.params[2:]
        tgt = _direct_targets(ctx.params)
        tg = tgt[2:] if (tgt is not None and _grad_mode() == "nct") else None
        arena = None if tg is not None else HL.ZeroArena(sum(_block_wgrad_floats(b) for b in (B1, B2, B3)), x16.device)
        T = (lambda i, j: None) if tg is None else (lambda i, j: tuple(tg[i:j]) + ((None,) if j - i == 4 else ()))
        g_x3, dW30, db30, dW31, db31, dWsc3 = _block_bwd(g3, g_d3pre, x3, h3, g7, B3, "e3", arena=arena, tg=T(9, 14))
        g_o2, g_d2pre = _avgpool_bwd(g_x3, d2, g14, B2.cout, 0.1)
        g_x2, dW20, db20, dW21, db21, dWsc2 = _block_bwd(g_o2, g_d2pre, x2, h2, g14, B2, "e2", arena=arena, tg=T(4, 9), hm=hm2)
        g_o1, g_d1pre = _avgpool_bwd(g_x2, d1, g28, B1.cout, 0.1)
        g_a0, dW10, db10, dW11, db11, _ = _block_bwd(g_o1, g_d1pre, a0, h1, g28, B1, "e1", arena=arena, tg=T(0, 4), hm=hm1)
        # conv_img (3 -> 64): weight gradient with the operand roles swapped (the 16-channel image is the N side),
        # dWs[t, c, n] = sum_p x[p + off_t, c] * g_a0[p, n]
        dWs = HL.wgrad(g_a0, x16, 9, [-o for o in g28.taps3x3()], g28.P, tag="e.img")       # [9, 16, 64]
        dWi = dWs.view(3, 3, 16, dWs.shape[2]).permute(3, 2, 0, 1)[:, : ctx.cin_img]
        dbi = _colsum(g_a0, g28.P, dWs.shape[2])
        if tg is not None:   # the block gradients were accumulated in place by the kernels
            return (None, dWi, dbi) + (None,) * 14
        if tgt is not None:  # one launch adds all sixteen buffers into the .grad tensors (the image conv's is role-swapped)
            blocks = (dW10, db10, dW11, db11, dW20, db20, dW21, db21, dWsc2, dW30, db30, dW31, db31, dWsc3)
            HL.unpack_wgrad_add([(dWs, tgt[0], True), (dbi, tgt[1])] + list(zip(blocks, tgt[2:])))
            return (None,) * 17
        u = HL.unpack_conv_wgrad
        return (None, dWi, dbi, u(dW10, 3, 3), db10, u(dW11, 3, 3), db11, u(dW20, 3, 3), db20, u(dW21, 3, 3), db21,
                u(dWsc2, 1, 1), u(dW30, 3, 3), db30, u(dW31, 3, 3), db31, u(dWsc3, 1, 1))


class FcFromHaloFn(torch.autograd.Function):
    """The two Linear heads (mu, log-variance) on the flattened 7x7x256 activation kept in halo order: the weight columns are
    gathered into halo order (zeros at halo positions).  Native GEMMs (mv_gemm), fp32 outputs."""

    @staticmethod
    def forward(ctx, h, w_mu, b_mu, w_lv, b_lv, s0=7):
        from .linear_native import gemm
        ch = h.shape[1]
        S = (s0 + 1) * (s0 + 1)
        gather, scatter = _fc_perm(h.device, s0, ch)
        n_img = h.shape[0] // S
        Fp, k_mu, k_lv = gather.numel(), w_mu.shape[0], w_lv.shape[0]
        wp = torch.empty(k_mu + k_lv, Fp, device=h.device, dtype=torch.bfloat16)                       # [2L, S*ch], halo column order
        for w, r0 in ((w_mu, 0), (w_lv, k_mu)):
            w = w.detach().float()
            C.check(C.lib().mv_gather_cast(w.data_ptr(), w.shape[0], w.shape[1], w.stride(0), gather.data_ptr(), w.shape[0], Fp, 1,
                                           wp[r0:].data_ptr(), Fp, C.stream()), "mv_gather_cast")
        h2 = h.reshape(n_img, S * ch)
        N = wp.shape[0]
        out = torch.empty(n_img, N, device=h.device, dtype=torch.float32)
        gemm(h2, wp, n_img, N, S * ch, out, bias=torch.cat([b_mu.detach(), b_lv.detach()]).float().contiguous(), out_kind=1, tag="enc.fc")
        ctx.save_for_backward(h2, wp)
        ctx.scatter, ctx.split, ctx.ch = scatter, w_mu.shape[0], ch
        return out[:, : w_mu.shape[0]].contiguous(), out[:, w_mu.shape[0]:].contiguous()

    @staticmethod
    def backward(ctx, g_mu, g_lv):
        from .linear_native import _to_bf16_padded, gemm
        h2, wp = ctx.saved_tensors
        n, F_ = h2.shape
        N = wp.shape[0]
        g = torch.cat([g_mu, g_lv], 1)
        gb = g.sum(0, dtype=torch.float32)
        g16 = _to_bf16_padded(g)
        gw = torch.zeros(N, F_, device=g.device, dtype=torch.float32)
        gemm(g16, h2, N, F_, n, gw, a_mn=True, b_mn=True, out_kind=2, tag="enc.fc.w")
        gh = torch.empty(n, F_, device=g.device, dtype=torch.bfloat16)
        gemm(g16, wp, n, F_, N, gh, b_mn=True, tag="enc.fc.d")
        k, Fr = ctx.split, ctx.scatter.numel()
        gw_mu = _gather_f32(gw[:k], F_, ctx.scatter, 1, k, Fr)
        gw_lv = _gather_f32(gw[k:], F_, ctx.scatter, 1, N - k, Fr)
        return gh.reshape(-1, ctx.ch), gw_mu, gb[:k], gw_lv, gb[k:], None


def encoder_params(enc, tag):
    r = getattr(enc, f"resnet_{tag}")
    b1, b2, b3 = r[0], r[2], r[4]
    ci = getattr(enc, f"conv_img_{tag}")
    return (ci.weight, ci.bias,
            b1.conv_layers[0].weight, b1.conv_layers[0].bias, b1.conv_layers[2].weight, b1.conv_layers[2].bias,
            b2.conv_layers[0].weight, b2.conv_layers[0].bias, b2.conv_layers[2].weight, b2.conv_layers[2].bias,
            b2.shortcut_layer.weight,
            b3.conv_layers[0].weight, b3.conv_layers[0].bias, b3.conv_layers[2].weight, b3.conv_layers[2].bias,
            b3.shortcut_layer.weight)


def encoder_forward(enc, x):
    out = ModelOutput()
    for tag, keys in (("u", ("embedding", "log_covariance")), ("w", ("style_embedding", "style_log_covariance"))):
        if tag == "w" and not enc.multiple_latent:
            continue
        h = EncoderStackFn.apply(x, *encoder_params(enc, tag))
        mu, lv = FcFromHaloFn.apply(h, getattr(enc, f"fc_mu_{tag}").weight, getattr(enc, f"fc_mu_{tag}").bias,
                                    getattr(enc, f"fc_lv_{tag}").weight, getattr(enc, f"fc_lv_{tag}").bias, 7)
        out[keys[0]], out[keys[1]] = mu, lv
    return out
