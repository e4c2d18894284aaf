"""Native (tcgen05) execution of the PolyMNIST ResNet decoder (reference: models/nn/mmnist.py:214-251,321-366).

`DecoderResnetMMNIST.forward` routes here on CUDA when the model computes in bf16.  The whole convolutional
stack (3 ResnetBlocks, 2 nearest upsamplings, the image head) runs as hand-written sm_100a kernels through
the C-ABI — mv_tapgemm (forward + data gradients), mv_wgrad (weight gradients) and the HBM-bound helpers
of csrc/halo_ops.cu — on activations kept in the shared-halo NHWC bf16 layout; fp32 master weights are
packed to bf16 tap-major matrices once per step.  The fully connected input layer (1.6 of 213 MFLOP per
image) is a library GEMM whose permuted weight makes it emit the halo layout directly.

There is no fallback inside this path: if the C-ABI library is missing, it raises.
"""
import torch
import torch.nn.functional as F

from .. import _cabi as C
from ..containers import ModelOutput
from . import functional as NF
from . import halo as HL

_LRELU = 0.2


def use_native(x):
    """Native path: CUDA tensor, backend not forced to torch, bf16 compute (autocast active)."""
    if not x.is_cuda or NF.backend() == "torch":
        return False
    return NF.backend() == "native" or torch.is_autocast_enabled()


def status(which):
    if which == "decoder":
        return "native sm_100a: tcgen05 tap-GEMM conv fwd/dgrad + tcgen05 wgrad + halo kernels (fc: cuBLAS)"
    return "torch (cuDNN/cuBLAS library calls)"


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
    C.check(C.lib().mv_colsum(G.data_ptr(), P, G.stride(0), n, out.data_ptr(), C.stream()), "mv_colsum")
    return out


class _Block:
    """bf16 packs of one ResnetBlock's weights (forward and data-gradient orientation)."""

    def __init__(self, w0, b0, w1, b1, wsc):
        self.cin, self.hid, self.cout = w0.shape[1], w0.shape[0], w1.shape[0]
        self.w0, self.w0d = HL.pack_conv_weight(w0), HL.pack_conv_weight_dgrad(w0)
        self.w1, self.w1d = HL.pack_conv_weight(w1), HL.pack_conv_weight_dgrad(w1)
        self.b0, self.b1 = b0.detach().float().contiguous(), b1.detach().float().contiguous()
        self.wsc = self.wscd = None
        if wsc is not None:
            self.wsc, self.wscd = HL.pack_conv_weight(wsc), HL.pack_conv_weight_dgrad(wsc)


def _block_fwd(x, g, blk, tag):
    """x_s + 0.1 * lrelu(conv1(lrelu(conv0(x))))  ->  (out, saved h, saved d)."""
    taps = g.taps3x3()
    xs = x if blk.wsc is None else HL.tapgemm(x, blk.wsc, 1, [0], blk.cout, g.P, geom=g, tag=f"{tag}.sc")
    h = HL.tapgemm(x, blk.w0, 9, taps, blk.hid, g.P, bias=blk.b0, act="lrelu", geom=g, tag=f"{tag}.c0")
    d = torch.empty(g.P, blk.cout, device=x.device, dtype=torch.bfloat16)
    out = HL.tapgemm(h, blk.w1, 9, taps, blk.cout, g.P, bias=blk.b1, act="lrelu", alpha=0.1, res=xs, out2=d, out2_pre=True,
                     geom=g, tag=f"{tag}.c1")
    return out, h, d


def _block_bwd(g_out, g_dpre, x, h, g, blk, tag, need_gx=True):
    """g_out: gradient of the block output; g_dpre = 0.1 * g_out * lrelu'(d) (produced upstream).
    Returns (g_x, dW0, db0, dW1, db1, dWsc)."""
    taps = g.taps3x3()
    dW1 = HL.wgrad(h, g_dpre, 9, taps, g.P, tag=f"{tag}.c1")
    db1 = _colsum(g_dpre, g.P, blk.cout)
    g_hpre = HL.tapgemm(g_dpre, blk.w1d, 9, taps, blk.hid, g.P, dact1=h, slope1=_LRELU, geom=g, tag=f"{tag}.c1d")
    dW0 = HL.wgrad(x, g_hpre, 9, taps, g.P, tag=f"{tag}.c0")
    db0 = _colsum(g_hpre, g.P, blk.hid)
    dWsc = None
    g_short = g_out
    if blk.wsc is not None:
        dWsc = HL.wgrad(x, g_out, 1, [0], g.P, tag=f"{tag}.sc")
        g_short = HL.tapgemm(g_out, blk.wscd, 1, [0], blk.cin, g.P, geom=g, tag=f"{tag}.scd") if need_gx else None
    g_x = None
    if need_gx:
        g_x = HL.tapgemm(g_hpre, blk.w0d, 9, taps, blk.cin, g.P, res=g_short, geom=g, tag=f"{tag}.c0d")
    return g_x, dW0, db0, dW1, db1, dWsc


class DecoderStackFn(torch.autograd.Function):
    """h0 (halo matrix at 7x7, 256 ch, bf16) + the 17 conv parameters -> reconstruction [n_img, 3, 28, 28] bf16."""

    @staticmethod
    def forward(ctx, h0, n_img, *params):
        (w10, b10, w11, b11, wsc1, w20, b20, w21, b21, wsc2, w30, b30, w31, b31, wh, bh) = params
        dev = h0.device
        B1 = _Block(w10, b10, w11, b11, wsc1)
        B2 = _Block(w20, b20, w21, b21, wsc2)
        B3 = _Block(w30, b30, w31, b31, None)
        whp = torch.zeros(16, wh.shape[1], 3, 3, device=dev, dtype=wh.dtype)
        whp[: wh.shape[0]] = wh.detach()
        bhp = torch.zeros(16, device=dev, dtype=torch.float32)
        bhp[: wh.shape[0]] = bh.detach().float()
        g7 = HL.Geom(n_img, 7, 7)
        o1, h1, d1 = _block_fwd(h0, g7, B1, "b1")
        u1, g14 = _upsample_fwd(o1, g7, B1.cout)
        o2, h2, d2 = _block_fwd(u1, g14, B2, "b2")
        u2, g28 = _upsample_fwd(o2, g14, B2.cout)
        o3, h3, d3 = _block_fwd(u2, g28, B3, "b3")
        n_ch = wh.shape[0]
        recon = torch.empty(n_img, n_ch, 28, 28, device=dev, dtype=torch.bfloat16)
        HL.tapgemm(o3, HL.pack_conv_weight(whp), 9, g28.taps3x3(), 16, g28.P, bias=bhp, act="lrelu", geom=g28, nchw_out=recon,
                   n_valid=n_ch, tag="head")
        ctx.save_for_backward(h0, h1, d1, u1, h2, d2, u2, h3, d3, o3, recon)
        ctx.packs = (B1, B2, B3, HL.pack_conv_weight_dgrad(whp))
        ctx.n_img, ctx.n_ch = n_img, n_ch
        return recon

    @staticmethod
    def backward(ctx, g_recon):
        h0, h1, d1, u1, h2, d2, u2, h3, d3, o3, recon = ctx.saved_tensors
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
        dWh = HL.wgrad(o3, gh, 9, g28.taps3x3(), g28.P, tag="head")
        dbh = _colsum(gh, g28.P, 16)
        g_dpre3 = torch.empty(g28.P, B3.cout, device=h0.device, dtype=torch.bfloat16)
        g_o3 = HL.tapgemm(gh, whd, 9, g28.taps3x3(), B3.cout, g28.P, out2=g_dpre3, alpha2=0.1, dact2=d3, slope2=_LRELU, geom=g28,
                          tag="head.d")
        g_u2, dW30, db30, dW31, db31, _ = _block_bwd(g_o3, g_dpre3, u2, h3, g28, B3, "b3")
        g_o2, g_dpre2 = _upsample_bwd(g_u2, d2, g14, B2.cout, 0.1)
        g_u1, dW20, db20, dW21, db21, dWsc2 = _block_bwd(g_o2, g_dpre2, u1, h2, g14, B2, "b2")
        g_o1, g_dpre1 = _upsample_bwd(g_u1, d1, g7, B1.cout, 0.1)
        need_h0 = ctx.needs_input_grad[0]
        g_h0, dW10, db10, dW11, db11, dWsc1 = _block_bwd(g_o1, g_dpre1, h0, h1, g7, B1, "b1", need_gx=need_h0)
        if g_h0 is not None:
            g_h0 = g_h0[: h0.shape[0]]
        u = HL.unpack_conv_wgrad
        grads = (u(dW10, 3, 3), db10, u(dW11, 3, 3), db11, u(dWsc1, 1, 1), u(dW20, 3, 3), db20, u(dW21, 3, 3), db21,
                 u(dWsc2, 1, 1), u(dW30, 3, 3), db30, u(dW31, 3, 3), db31, u(dWh, 3, 3)[:n_ch], dbh[:n_ch])
        return (g_h0, None) + grads


_PERM_CACHE = {}


def _fc_perm(device):
    """(gather, scatter): `gather` [64*256] maps halo feature (y'*8 + x)*256 + c to fc row c*49 + (y'-1)*7 + x (halo
    positions -> the extra all-zero row 12544); `scatter` [12544] is its inverse on the real rows."""
    key = str(device)
    if key not in _PERM_CACHE:
        idx = torch.full((8, 8, 256), 256 * 49, dtype=torch.long)
        c = torch.arange(256)
        for y in range(1, 8):
            for x in range(7):
                idx[y, x] = c * 49 + (y - 1) * 7 + x
        idx = idx.reshape(-1)
        inv = torch.empty(256 * 49, dtype=torch.long)
        pos = torch.nonzero(idx < 256 * 49).squeeze(1)
        inv[idx[pos]] = pos
        _PERM_CACHE[key] = (idx.to(device), inv.to(device))
    return _PERM_CACHE[key]


class FcHaloFn(torch.autograd.Function):
    """fc of the decoder emitting the 7x7 halo matrix directly: rows of the weight are gathered into halo order
    (zero rows at halo positions), so `z @ Wp^T + bp` IS the [n_img*64, 256] activation matrix.  Library GEMMs
    (1.6 of 213 MFLOP per image); the backward gathers the real rows back (no scatter-add)."""

    @staticmethod
    def forward(ctx, z, weight, bias):
        gather, scatter = _fc_perm(z.device)
        w_ext = torch.cat([weight.detach(), weight.new_zeros(1, weight.shape[1])], 0).to(torch.bfloat16)
        b_ext = torch.cat([bias.detach(), bias.new_zeros(1)], 0).to(torch.bfloat16)
        wp = w_ext[gather]
        zb = z.detach().to(torch.bfloat16)
        h0 = torch.addmm(b_ext[gather], zb, wp.t())
        ctx.save_for_backward(zb, wp)
        ctx.scatter = scatter
        return h0

    @staticmethod
    def backward(ctx, g):
        zb, wp = ctx.saved_tensors
        g = g.reshape(zb.shape[0], -1).to(torch.bfloat16)
        gz = (g @ wp).float() if ctx.needs_input_grad[0] else None
        gw = (g.t() @ zb).float()[ctx.scatter]
        gb = g.float().sum(0)[ctx.scatter]
        return gz, gw, gb


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
    h0 = FcHaloFn.apply(zz, dec.fc.weight, dec.fc.bias).view(n_img * 64, 256)  # halo matrix at 7x7
    recon = DecoderStackFn.apply(h0, n_img, *decoder_params(dec))
    recon = recon.view(*z.size()[:-1], *recon.shape[1:])
    return ModelOutput(reconstruction=recon)


def encoder_forward(enc, x):
    raise NotImplementedError("the ResNet encoder runs on the library path in this round")
