"""Native (tcgen05) execution of the fully connected stacks (reference: models/nn/default_architectures.py:21-130,225-258 and the
nn.Linear layers of mmnist.py / svhn.py): a whole chain  x -> act_0(x W_0^T + b_0) -> ... -> act_{n-1}(. W_{n-1}^T + b_{n-1})
runs as one autograd Function over the general GEMM of csrc/gemm.cu (mv_gemm) through the C-ABI:

  forward    one GEMM per layer with bias + ReLU / Sigmoid fused in the epilogue (bf16 activations, fp32 accumulate);
  backward   per layer: weight gradient (both operands MN-major, fp32 reductions straight into the parameter's .grad when the
             trainer opted in), bias gradient (column sums), data gradient with the ReLU derivative of the layer below fused
             into its epilogue; the Sigmoid derivative of the last layer is one elementwise kernel (mv_act_bwd).

fp32 master weights are packed to bf16 (K padded to a multiple of 8 for the 16-byte TMA pitch) in ONE launch per chain.
No fallback: a missing C-ABI library raises."""
import torch

from .. import _cabi as C
from . import halo as HL
from . import resnet_native as RN

_ACT_CODE = {"none": 0, "relu": 1, "lrelu": 2, "sigmoid": 3}


def _pad8(n):
    return (n + 7) // 8 * 8


def gemm(A, B, M, N, K, out, *, a_mn=False, b_mn=False, bias=None, act="none", alpha=1.0, dact=None, dslope=0.0, out_kind=0, tag=None):
    """out[m, n] (+)= epilogue(alpha * sum_k A(m, k) B(n, k)); see mv_gemm in include/multivae_b200.h."""
    a = C.GemmArgs()
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.stride(-1) == 1 and B.stride(-1) == 1
    a.A, a.M, a.a_ld, a.a_mn = A.data_ptr(), M, A.stride(0), int(a_mn)
    a.B, a.N, a.b_ld, a.b_mn = B.data_ptr(), N, B.stride(0), int(b_mn)
    a.K = K
    a.bias = None if bias is None else C.ptr(bias)
    a.act, a.alpha = _ACT_CODE[act], float(alpha)
    if dact is not None:
        assert dact.dtype == torch.bfloat16 and dact.stride(-1) == 1
        a.dact, a.dact_ld, a.dslope = dact.data_ptr(), dact.stride(0), float(dslope)
    assert out.stride(-1) == 1 and out.dtype == (torch.bfloat16 if out_kind == 0 else torch.float32)
    a.out, a.out_ld, a.out_kind = out.data_ptr(), out.stride(0), out_kind
    kw = {}
    if tag is not None:   # algorithmic work of this launch for bench.py's roofline (flops; operand + result bytes)
        ob = M * N * (2 if out_kind == 0 else (4 if out_kind == 1 else 8))
        kw["tag"] = f"{tag}|f={2 * M * N * K}|b={2 * (M * K + N * K) + ob}"
    C.check(C.lib().mv_gemm(a, C.stream(), **kw), "mv_gemm")
    return out


def _to_bf16_padded(x):
    """[B, K] any float dtype -> contiguous bf16 [B, K8] (K8 = K rounded up to 8, zero columns), as a [B, K] view of it."""
    B, K = x.shape
    K8 = _pad8(K)
    if K8 == K:
        return x.detach().to(torch.bfloat16).contiguous()
    buf = torch.zeros(B, K8, device=x.device, dtype=torch.bfloat16)
    buf[:, :K].copy_(x.detach())
    return buf[:, :K]


def _pack_weights(weights):
    """fp32 [N_i, K_i] masters (lists = heads concatenated along N) -> bf16 [N, K8] operands, ONE launch for the chain."""
    specs, outs = [], []
    for w in weights:
        group = w if isinstance(w, (list, tuple)) else [w]
        K = group[0].shape[1]
        N = sum(g.shape[0] for g in group)
        buf = torch.empty(N, _pad8(K), device=group[0].device, dtype=torch.bfloat16)
        outs.append(buf)
        r = 0
        for g in group:
            specs.append((g, buf[r:r + g.shape[0]]))
            r += g.shape[0]
    items = (C.PackItem * len(specs))()
    for it, (w, dst) in zip(items, specs):
        w = w.detach()
        assert w.dtype == torch.float32 and w.is_contiguous() and w.is_cuda and w.dim() == 2
        it.src, it.dst_fwd, it.dst_dgrad = w.data_ptr(), dst.data_ptr(), None
        it.N, it.C, it.T, it.Npad, it.Cpad = w.shape[0], w.shape[1], 1, w.shape[0], dst.shape[1]
    for i0 in range(0, len(specs), 24):   # MV_PACK_MAX_ITEMS per launch
        n = min(24, len(specs) - i0)
        sub = (C.PackItem * n)(*[items[i0 + j] for j in range(n)])
        C.check(C.lib().mv_pack_tc(sub, n, C.stream()), "mv_pack_tc")   # T = 1: a cast (+ zero padding of K to a multiple of 8)
    return outs


class MLPChainFn(torch.autograd.Function):
    """y = chain of Linear layers with fused activations.  forward(ctx, x, spec, *params): spec = (acts, groups, out_fp32) with
    acts[i] in {"none", "relu", "sigmoid"} and groups[i] = number of Linear heads concatenated in layer i (only the last layer may
    have several); params = (W, b) per head in layer order.  Returns [B, N_last] (fp32 when out_fp32 else bf16)."""

    @staticmethod
    def forward(ctx, x, spec, *params):
        acts, groups, out_fp32 = spec
        layers, i = [], 0
        for g in groups:
            layers.append([(params[i + 2 * j], params[i + 2 * j + 1]) for j in range(g)])
            i += 2 * g
        x2 = x.reshape(-1, x.shape[-1])
        Bt = x2.shape[0]
        dev = x.device
        packs = _pack_weights([[w for w, _ in l] for l in layers])
        h = _to_bf16_padded(x2)
        saved_in = []
        last_buf = None
        for li, (l, wb) in enumerate(zip(layers, packs)):
            N, K = wb.shape[0], l[0][0].shape[1]
            bias = l[0][1].detach().float() if len(l) == 1 else torch.cat([b.detach().float() for _, b in l])
            last = li == len(layers) - 1
            saved_in.append(h)
            if last and out_fp32:
                out = torch.empty(Bt, N, device=dev, dtype=torch.float32)
                gemm(h, wb, Bt, N, K, out, bias=bias.contiguous(), act=acts[li], out_kind=1, tag=f"fc{li}")
            else:
                N8 = _pad8(N)
                # ragged widths: the padding columns are never written by the clipped TMA store; zero them once so that whatever
                # reads the whole [Bt, N8] buffer (the activation derivative of the last layer) sees finite values
                buf = torch.empty(Bt, N8, device=dev, dtype=torch.bfloat16) if N8 == N else torch.zeros(Bt, N8, device=dev, dtype=torch.bfloat16)
                out = buf[:, :N]
                last_buf = buf
                gemm(h, wb, Bt, N, K, out, bias=bias.contiguous(), act=acts[li], out_kind=0, tag=f"fc{li}")
            h = out
        ctx.save_for_backward(*saved_in, *packs, h, last_buf if last_buf is not None else h)
        ctx.layers_meta = [[(w.shape[0], w.shape[1]) for w, _ in l] for l in layers]
        ctx.params = params
        ctx.spec = spec
        ctx.n_layers = len(layers)
        ctx.x_shape = x.shape
        return h.reshape(*x.shape[:-1], h.shape[-1])

    @staticmethod
    def backward(ctx, g):
        acts, groups, out_fp32 = ctx.spec
        n = ctx.n_layers
        saved = ctx.saved_tensors
        ins, packs, y, ybuf = saved[:n], saved[n:2 * n], saved[2 * n], saved[2 * n + 1]
        lib = C.lib()
        dev = g.device
        g2 = g.reshape(-1, g.shape[-1])
        Bt, Nl = g2.shape
        targets = RN._direct_targets(ctx.params)   # .grad tensors when the trainer opted in, else None
        # gradient w.r.t. the last layer's pre-activation, as the bf16 GEMM operand
        if acts[-1] in ("sigmoid", "relu", "lrelu"):
            N8 = _pad8(Nl)
            if N8 == Nl:
                gc, yb = g2.contiguous(), y
            else:   # ragged width: both operands as whole [Bt, N8] buffers (zero padding columns on both sides)
                gc = torch.zeros(Bt, N8, device=dev, dtype=g2.dtype)
                gc[:, :Nl].copy_(g2)
                yb = ybuf   # the whole [Bt, N8] buffer behind y (zero padding columns)
                assert yb.shape == (Bt, N8) and yb.is_contiguous()
            dbuf = torch.empty(Bt, N8, device=dev, dtype=torch.bfloat16)
            C.check(lib.mv_act_bwd(C.ptr(gc), C.dtype_code(gc), C.ptr(yb), C.ptr(dbuf), Bt * N8, _ACT_CODE[acts[-1]],
                                   0.2 if acts[-1] == "lrelu" else 0.0, C.stream()), "mv_act_bwd")
            d = dbuf[:, :Nl]
        else:
            d = _to_bf16_padded(g2)
        grads = [None] * len(ctx.params)
        arena_n = 0
        if targets is None:
            for lm in ctx.layers_meta:
                arena_n += sum(nn * kk + _pad8(nn) + 8 for nn, kk in lm)
            arena = HL.ZeroArena(arena_n + 64, dev)
        pi_end = len(ctx.params)
        g_x = None
        for li in range(n - 1, -1, -1):
            lm = ctx.layers_meta[li]
            N = sum(nn for nn, _ in lm)
            K = lm[0][1]
            x_in, wb = ins[li], packs[li]
            pi0 = pi_end - 2 * len(lm)
            # weight and bias gradients, head by head (column slices of d)
            r = 0
            dp = d if d.shape[1] % 8 == 0 else None
            for j, (nn, kk) in enumerate(lm):
                dj = d[:, r:r + nn]
                if targets is not None:
                    dW, db = targets[pi0 + 2 * j], targets[pi0 + 2 * j + 1]
                else:
                    dW, db = arena.take(nn, kk), arena.take(_pad8(nn))[:nn]
                    grads[pi0 + 2 * j], grads[pi0 + 2 * j + 1] = dW, db
                if (dj.data_ptr() % 16) or (dj.stride(0) % 8):   # unaligned head slice: compact copy
                    dj = _to_bf16_padded(dj)
                gemm(dj, x_in, nn, kk, Bt, dW, a_mn=True, b_mn=True, out_kind=2, tag=f"fc{li}.w")
                if nn % 8 == 0:
                    C.check(lib.mv_colsum_any(dj.data_ptr(), Bt, dj.stride(0), nn, db.data_ptr(), C.stream()), "mv_colsum_any")
                else:   # narrow head (e.g. 20 latent dims): tiny library reduction
                    db.add_(dj.float().sum(0))
                r += nn
            pi_end = pi0
            # data gradient (with the ReLU derivative of the layer below fused into the epilogue)
            if li > 0:
                assert acts[li - 1] in ("relu", "lrelu", "none")
                K8 = _pad8(K)
                dx = torch.empty(Bt, K8, device=dev, dtype=torch.bfloat16)[:, :K]
                gemm(d, wb, Bt, K, N, dx, b_mn=True, dact=None if acts[li - 1] == "none" else x_in,
                     dslope=0.2 if acts[li - 1] == "lrelu" else 0.0, out_kind=0, tag=f"fc{li}.d")
                d = dx
            elif ctx.needs_input_grad[0]:
                g_x = torch.empty(Bt, K, device=dev, dtype=torch.float32)
                gemm(d, wb, Bt, K, N, g_x, b_mn=True, out_kind=1, tag=f"fc{li}.d")
                g_x = g_x.reshape(ctx.x_shape)
        return (g_x, None) + tuple(grads)


def mlp_chain(x, layers, acts, out_fp32=False):
    """layers: list of nn.Linear or lists of nn.Linear (heads sharing the input, last layer only)."""
    groups, params = [], []
    for l in layers:
        heads = l if isinstance(l, (list, tuple)) else [l]
        groups.append(len(heads))
        for h in heads:
            params += [h.weight, h.bias]
    return MLPChainFn.apply(x, (tuple(acts), tuple(groups), out_fp32), *params)
