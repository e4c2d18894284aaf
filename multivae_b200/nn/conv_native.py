"""Native execution of the small strided-convolution networks (reference: models/nn/svhn.py:7-70 Encoder/Decoder_VAE_SVHN,
models/nn/mmnist.py:78-110,173-207 EncoderConvMMNIST_adapted / DecoderConvMMNIST): every contraction runs on the general tcgen05
GEMM (csrc/gemm.cu) around the gather kernels of csrc/im2col.cu — no cuDNN, no cuBLAS.

  nn.Conv2d (k, stride 2)         cols = im2col(x);  y = relu(cols W^T + b)         [GEMM epilogue: bias + ReLU]
      weight gradient             dW += dY^T cols                                   [both operands MN-major, into weight.grad]
      data gradient               dcols = dY W;  dx = col2im(dcols) * relu'(x)      [mask fused into the gather]
  nn.ConvTranspose2d              cols = x W;  y = act(col2im(cols) + b)            [bias + ReLU / Sigmoid fused into the gather]
      weight gradient             dW += x^T im2col(dY)
      data gradient               dx = im2col(dY) W^T * relu'(x)                    [mask fused into the GEMM epilogue]

Activations are dense NHWC bf16 matrices [n * H * W, C]; the patch matrices use tap-major columns (t*C + c: the channels of a
pixel stay contiguous, so the gathers move 16-byte vectors), weights are re-laid out to that order while they are cast to bf16
(mv_pack_tc, one launch per stack) and weight gradients return to the torch layout in one launch (mv_unpack_tc_add).
These networks are < 10 MFLOP per sample: the step is bound by launches and HBM, not by the tensor pipe."""
import torch

from .. import _cabi as C
from . import halo as HL
from . import resnet_native as RN
from .linear_native import _pack_weights, _pad8, _to_bf16_padded, gemm

_ACT = {"none": 0, "relu": 1, "lrelu": 2, "sigmoid": 3}


def _geom(n_img, H, W, Cc, nchw, k, s, p, gh, gw, ld, tc=True):
    g = C.ConvGeom()
    g.n_img, g.H, g.W, g.C, g.nchw = n_img, H, W, Cc, int(nchw)
    g.kh, g.kw, g.stride, g.pad, g.grid_h, g.grid_w, g.ld, g.tc_order = k, k, s, p, gh, gw, ld, int(tc)
    return g


def _tc_items(pairs):
    """(fp32 [N, C, kh, kw] or [N, C, T] tensor, packed [N, ld] tensor) pairs -> PackItem array for mv_pack_tc / mv_unpack_tc_add."""
    items = (C.PackItem * len(pairs))()
    for it, (w, packed) in zip(items, pairs):
        assert w.dtype == torch.float32 and w.is_contiguous() and packed.is_contiguous()
        it.src, it.dst_fwd, it.dst_dgrad = w.data_ptr(), packed.data_ptr(), None
        it.N, it.C, it.T = w.shape[0], w.shape[1], w[0, 0].numel()
        it.Npad, it.Cpad = w.shape[0], packed.shape[1]
    return items


def pack_tc(weights):
    """fp32 conv weights [N, C, kh, kw] -> bf16 [N, pad8(T*C)] with tap-major columns, all in one launch."""
    outs = [torch.empty(w.shape[0], _pad8(w[0].numel()), device=w.device, dtype=torch.bfloat16) for w in weights]
    C.check(C.lib().mv_pack_tc(_tc_items([(w.detach(), o) for w, o in zip(weights, outs)]), len(weights), C.stream()), "mv_pack_tc")
    return outs


def unpack_tc_add(pairs):
    """grad[n, c, t] += dW[n, t*C + c] for every (dW fp32 [N, ld], grad fp32 [N, C, kh, kw]) pair in one launch."""
    items = _tc_items([(g, d) for d, g in pairs])
    for it, (d, g) in zip(items, pairs):
        it.src, it.dst_fwd = d.data_ptr(), g.data_ptr()
    C.check(C.lib().mv_unpack_tc_add(items, len(pairs), C.stream()), "mv_unpack_tc_add")


def im2col(src, g):
    cols = torch.empty(g.n_img * g.grid_h * g.grid_w, g.ld, device=src.device, dtype=torch.bfloat16)
    C.check(C.lib().mv_im2col(C.ptr(src), C.dtype_code(src), C.ptr(cols), g, C.stream(),
                              tag=f"|b={cols.numel() * 2 + src.numel() * src.element_size()}"), "mv_im2col")
    return cols


def col2im(cols, g, bias=None, act="none", dact=None, dslope=0.0):
    shape = (g.n_img, g.C, g.H, g.W) if g.nchw else (g.n_img * g.H * g.W, g.C)
    dst = torch.empty(shape, device=cols.device, dtype=torch.bfloat16)
    C.check(C.lib().mv_col2im(C.ptr(cols), C.dtype_code(cols), C.ptr(dst), g, None if bias is None else C.ptr(bias), _ACT[act],
                              None if dact is None else C.ptr(dact), float(dslope), C.stream(),
                              tag=f"|b={cols.numel() * cols.element_size() + dst.numel() * (2 if dact is None else 4)}"), "mv_col2im")
    return dst


def _colsum(d, n, out):
    C.check(C.lib().mv_colsum_any(d.data_ptr(), d.shape[0], d.stride(0), n, out.data_ptr(), C.stream()), "mv_colsum_any")


def _grad_slots(ctx_params, dev, sizes):
    """(.grad targets | None, list of fp32 accumulators per parameter)."""
    targets = RN._direct_targets(ctx_params)
    if targets is not None:
        return targets, targets
    arena = HL.ZeroArena(sum(_pad8(n) + 8 for n in sizes) + 64, dev)
    return None, [arena.take(_pad8(n))[:n] for n in sizes]


class ConvEncoderFn(torch.autograd.Function):
    """x [B, C, H, W] fp32 -> [B, sum of head widths] fp32.  spec = ((k, stride, pad) per conv, n_heads); params = (w, b) of the
    convolutions then of the heads (Conv2d over the whole final feature map = Linear)."""

    @staticmethod
    def forward(ctx, x, spec, *params):
        convs, n_heads = spec
        nc = len(convs)
        cw = [(params[2 * i], params[2 * i + 1]) for i in range(nc)]
        hw = [(params[2 * nc + 2 * j], params[2 * nc + 2 * j + 1]) for j in range(n_heads)]
        B, Cc, H, W = x.shape
        x = x.detach().contiguous()
        packs = pack_tc([w for w, _ in cw] + [w for w, _ in hw])
        cur, nchw = x, True
        cols_l, acts_l, geoms = [], [], []
        for i, ((w, b), (k, s, p)) in enumerate(zip(cw, convs)):
            Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
            K, N = Cc * k * k, w.shape[0]
            g = _geom(B, H, W, Cc, nchw, k, s, p, Ho, Wo, _pad8(K))
            cols = im2col(cur, g)
            h = torch.empty(B * Ho * Wo, N, device=x.device, dtype=torch.bfloat16)
            gemm(cols, packs[i], B * Ho * Wo, N, K, h, bias=b.detach().float().contiguous(), act="relu", tag=f"conv{i}")
            cols_l.append(cols); acts_l.append(h); geoms.append(g)
            cur, nchw, H, W, Cc = h, False, Ho, Wo, N
        kh = hw[0][0].shape[2]
        assert kh == H and hw[0][0].shape[3] == W, "the heads must cover the whole final feature map"
        Kh = Cc * H * W
        assert Kh % 8 == 0
        gh = _geom(B, H, W, Cc, False, kh, 1, 0, 1, 1, Kh)
        cols_h = cur.view(B, Kh)   # tap-major columns of a full-map patch ARE the NHWC activation: no gather
        Wh = torch.cat(packs[nc:], 0)
        Nh = Wh.shape[0]
        out = torch.empty(B, Nh, device=x.device, dtype=torch.float32)
        gemm(cols_h, Wh, B, Nh, Kh, out, bias=torch.cat([b.detach().float() for _, b in hw]).contiguous(), out_kind=1, tag="heads")
        ctx.save_for_backward(*cols_l, *acts_l, Wh, *packs[:nc])
        ctx.geoms, ctx.gh, ctx.spec, ctx.params = geoms, gh, spec, params
        ctx.head_widths = [w.shape[0] for w, _ in hw]
        return out

    @staticmethod
    def backward(ctx, g_out):
        convs, n_heads = ctx.spec
        nc = len(convs)
        sv = ctx.saved_tensors
        cols_l, acts_l, Wh, packs = sv[:nc], sv[nc:2 * nc], sv[2 * nc], sv[2 * nc + 1:]
        dev = g_out.device
        params = ctx.params
        targets, acc = _grad_slots(params, dev, [p.numel() for p in params])
        B = g_out.shape[0]
        d = _to_bf16_padded(g_out)
        gh = ctx.gh
        Kh = gh.ld
        cols_h = acts_l[nc - 1].view(B, Kh)
        # weight gradients are produced in the packed (tap-major) layout and handed over in ONE launch at the end
        wsz = [params[2 * i].numel() for i in range(nc + n_heads)]
        warena = HL.ZeroArena(sum(wsz) + 8 * len(wsz), dev)
        dWp = [warena.take(params[2 * i].shape[0], params[2 * i][0].numel()) for i in range(nc + n_heads)]
        r = 0
        for j, n in enumerate(ctx.head_widths):
            dj = d[:, r:r + n]
            if (dj.data_ptr() % 16) or (dj.stride(0) % 8):
                dj = _to_bf16_padded(dj)
            gemm(dj, cols_h, n, Kh, B, dWp[nc + j], a_mn=True, b_mn=True, out_kind=2, tag="heads.w")
            if n % 8 == 0:
                _colsum(dj, n, acc[2 * nc + 2 * j + 1])
            else:
                acc[2 * nc + 2 * j + 1].add_(dj.float().sum(0))
            r += n
        dcols = torch.empty(B, Kh, device=dev, dtype=torch.float32)   # fp32: the gather below sums unrounded partial products
        gemm(d, Wh, B, Kh, d.shape[1], dcols, b_mn=True, out_kind=1, tag="heads.d")
        dpre = col2im(dcols, gh, dact=acts_l[nc - 1], dslope=0.0)
        for i in range(nc - 1, -1, -1):
            g = ctx.geoms[i]
            N, K = acts_l[i].shape[1], g.C * g.kh * g.kw
            P = dpre.shape[0]
            gemm(dpre, cols_l[i], N, K, P, dWp[i], a_mn=True, b_mn=True, out_kind=2, tag=f"conv{i}.w")
            _colsum(dpre, N, acc[2 * i + 1])
            if i > 0 or ctx.needs_input_grad[0]:
                dcols = torch.empty(P, g.ld, device=dev, dtype=torch.float32)
                gemm(dpre, packs[i], P, K, N, dcols[:, :K], b_mn=True, out_kind=1, tag=f"conv{i}.d")
                dpre = col2im(dcols, g, dact=acts_l[i - 1], dslope=0.0) if i > 0 else col2im(dcols, g)
        unpack_tc_add([(dWp[i], acc[2 * i].view(params[2 * i].shape)) for i in range(nc + n_heads)])
        g_x = dpre.float() if ctx.needs_input_grad[0] else None   # gradient of the input image (NCHW), only when asked for
        grads = (None,) * len(params) if targets is not None else tuple(a.view(p.shape) for a, p in zip(acc, params))
        return (g_x, None) + grads


class ConvTDecoderFn(torch.autograd.Function):
    """z [Bt, L] -> reconstruction [Bt, C, H, W] bf16.  spec = (first, ((k, stride, pad, out_pad) per transposed convolution),
    final_act, (C0, H0, W0)): `first` is "convt" (a ConvTranspose2d on the 1x1 latent, weight [L, C0, H0, W0], bias [C0]) or
    "linear" (nn.Linear(L, C0*H0*W0) + ReLU + Unflatten); hidden layers use ReLU.  params = (w, b) of the first layer then of the
    transposed convolutions."""

    @staticmethod
    def forward(ctx, z, spec, *params):
        first, convts, final_act, (C0, H0, W0) = spec
        nt = len(convts)
        w0, b0 = params[0], params[1]
        tw = [(params[2 + 2 * i], params[3 + 2 * i]) for i in range(nt)]
        Bt = z.shape[0]
        dev = z.device
        zb = _to_bf16_padded(z)
        packs = _pack_weights([w0.reshape(w0.shape[0], -1)]) + pack_tc([w for w, _ in tw])
        F0 = C0 * H0 * W0
        g0 = _geom(Bt, H0, W0, C0, False, H0, 1, 0, 1, 1, F0, tc=False)   # the first layer keeps the torch feature order (c, y, x)
        cols0 = torch.empty(Bt, F0, device=dev, dtype=torch.float32)
        if first == "convt":
            gemm(zb, packs[0], Bt, F0, z.shape[1], cols0, b_mn=True, out_kind=1, tag="dec0")
            x = col2im(cols0, g0, bias=b0.detach().float().contiguous(), act="relu")
        else:
            gemm(zb, packs[0], Bt, F0, z.shape[1], cols0, bias=b0.detach().float().contiguous(), act="relu", out_kind=1, tag="dec0")
            x = col2im(cols0, g0)
        xs, geoms = [x], []
        H, W, Cc = H0, W0, C0
        for i, ((w, b), (k, s, p, op)) in enumerate(zip(tw, convts)):
            Ho, Wo = (H - 1) * s - 2 * p + k + op, (W - 1) * s - 2 * p + k + op
            Co = w.shape[1]
            NT = Co * k * k
            last = i == nt - 1
            g = _geom(Bt, Ho, Wo, Co, last, k, s, p, H, W, _pad8(NT))
            cols = torch.empty(Bt * H * W, g.ld, device=dev, dtype=torch.float32)
            gemm(xs[-1], packs[1 + i], Bt * H * W, NT, Cc, cols[:, :NT], b_mn=True, out_kind=1, tag=f"convt{i}")
            y = col2im(cols, g, bias=b.detach().float().contiguous(), act=final_act if last else "relu")
            geoms.append(g)
            if not last:
                xs.append(y)
            H, W, Cc = Ho, Wo, Co
        ctx.save_for_backward(zb, *xs, y, *packs)
        ctx.geoms, ctx.g0, ctx.spec, ctx.params = geoms, g0, spec, params
        return y

    @staticmethod
    def backward(ctx, g):
        first, convts, final_act, (C0, H0, W0) = ctx.spec
        nt = len(convts)
        sv = ctx.saved_tensors
        zb, xs, y, packs = sv[0], sv[1:1 + nt], sv[1 + nt], sv[2 + nt:]
        params = ctx.params
        dev = g.device
        lib = C.lib()
        targets, acc = _grad_slots(params, dev, [p.numel() for p in params])
        Bt = zb.shape[0]
        gc = g.contiguous()
        if final_act == "none":
            dpre = gc.to(torch.bfloat16)
        else:
            dpre = torch.empty(y.shape, device=dev, dtype=torch.bfloat16)
            C.check(lib.mv_act_bwd(C.ptr(gc), C.dtype_code(gc), C.ptr(y), C.ptr(dpre), y.numel(), _ACT[final_act], 0.0, C.stream()), "mv_act_bwd")
        wsz = [params[2 + 2 * i].numel() for i in range(nt)]
        warena = HL.ZeroArena(sum(wsz) + 8 * nt, dev)
        dWp = [warena.take(params[2 + 2 * i].shape[0], params[2 + 2 * i][0].numel()) for i in range(nt)]
        for i in range(nt - 1, -1, -1):
            gm = ctx.geoms[i]
            Co, NT = gm.C, gm.C * gm.kh * gm.kw
            x_in = xs[i]
            P_in, Cin = x_in.shape
            if gm.nchw:
                C.check(lib.mv_chan_sum_nchw(C.ptr(dpre), Bt, Co, gm.H * gm.W, C.ptr(acc[3 + 2 * i]), C.stream()), "mv_chan_sum_nchw")
            else:
                _colsum(dpre, Co, acc[3 + 2 * i])
            dcols = im2col(dpre, gm)
            gemm(x_in, dcols, Cin, NT, P_in, dWp[i], a_mn=True, b_mn=True, out_kind=2, tag=f"convt{i}.w")
            dx = torch.empty(P_in, Cin, device=dev, dtype=torch.bfloat16)
            gemm(dcols, packs[1 + i], P_in, Cin, NT, dx, dact=x_in, dslope=0.0, tag=f"convt{i}.d")
            dpre = dx
        unpack_tc_add([(dWp[i], acc[2 + 2 * i].view(params[2 + 2 * i].shape)) for i in range(nt)])
        # first layer: dpre is the gradient of its pre-activation in NHWC [Bt * H0 * W0, C0]
        g0 = ctx.g0
        F0 = C0 * H0 * W0
        L = params[0].shape[0] if first == "convt" else params[0].shape[1]
        dcols0 = im2col(dpre, g0)
        g_z = None
        if first == "convt":
            _colsum(dpre, C0, acc[1])
            gemm(zb, dcols0, L, F0, Bt, acc[0].view(L, F0), a_mn=True, b_mn=True, out_kind=2, tag="dec0.w")
            if ctx.needs_input_grad[0]:
                g_z = torch.empty(Bt, L, device=dev, dtype=torch.float32)
                gemm(dcols0, packs[0], Bt, L, F0, g_z, out_kind=1, tag="dec0.d")
        else:
            _colsum(dcols0, F0, acc[1])
            gemm(dcols0, zb, F0, L, Bt, acc[0].view(F0, L), a_mn=True, b_mn=True, out_kind=2, tag="dec0.w")
            if ctx.needs_input_grad[0]:
                g_z = torch.empty(Bt, L, device=dev, dtype=torch.float32)
                gemm(dcols0, packs[0], Bt, L, F0, g_z, b_mn=True, out_kind=1, tag="dec0.d")
        grads = (None,) * len(params) if targets is not None else tuple(a.view(p.shape) for a, p in zip(acc, params))
        return (g_z, None) + grads


def conv_encoder(x, convs, heads):
    """convs: list of nn.Conv2d (each followed by ReLU); heads: list of nn.Conv2d over the whole final feature map."""
    spec = (tuple((c.kernel_size[0], c.stride[0], c.padding[0]) for c in convs), len(heads))
    params = []
    for c in list(convs) + list(heads):
        params += [c.weight, c.bias]
    out = ConvEncoderFn.apply(x, spec, *params)
    return torch.split(out, [h.weight.shape[0] for h in heads], dim=-1)


def convt_decoder(z, first, first_kind, convts, final_act, shape0):
    spec = (first_kind, tuple((c.kernel_size[0], c.stride[0], c.padding[0], c.output_padding[0]) for c in convts), final_act, tuple(shape0))
    params = [first.weight, first.bias]
    for c in convts:
        params += [c.weight, c.bias]
    return ConvTDecoderFn.apply(z, spec, *params)
