"""Shared-halo NHWC activation layout + thin Python wrappers over the tensor-core C-ABI entry points.

Layout (include/multivae_b200.h): a stage with images of H x W pixels and C channels is ONE bf16 matrix
[P, C], P = n_img*S + (W+1), S = (H+1)*(W+1): per image one zero row, then H rows of W pixels + 1 zero
pixel; one more zero row after the last image.  Every 3x3/pad-1 tap is a constant row shift of it."""
import torch

from .. import _cabi as C


class Geom:
    def __init__(self, n_img, H, W):
        self.n_img, self.H, self.W = n_img, H, W
        self.Wp = W + 1
        self.S = (H + 1) * self.Wp
        self.P = n_img * self.S + self.Wp

    def taps3x3(self):
        """Row offsets of the 9 taps in (kh, kw) order of a torch Conv2d weight."""
        return [(r - 1) * self.Wp + (s - 1) for r in range(3) for s in range(3)]

    def up2(self):
        return Geom(self.n_img, self.H * 2, self.W * 2)


def mask_rows(P):
    """Rows (64-bit words) of a sign-mask buffer for a P-row activation: whole 126-row tiles of the conv3 kernel."""
    return (P + 125) // 126 * 126


def to_halo(x, dtype=torch.bfloat16):
    """NCHW -> halo matrix [P, C] (torch ops; used for inputs and by the tests)."""
    n, c, h, w = x.shape
    g = Geom(n, h, w)
    buf = torch.zeros(g.P, c, device=x.device, dtype=dtype)
    v = buf[: n * g.S].view(n, h + 1, g.Wp, c)
    v[:, 1:, :w, :] = x.permute(0, 2, 3, 1).to(dtype)
    return buf, g


def from_halo(a, g):
    """halo matrix [P, C] -> NCHW (same dtype)."""
    c = a.shape[1]
    v = a[: g.n_img * g.S].view(g.n_img, g.H + 1, g.Wp, c)
    return v[:, 1:, : g.W, :].permute(0, 3, 1, 2).contiguous()


def pack_conv_weight(w, dtype=torch.bfloat16):
    """torch Conv2d weight [N, C, kh, kw] -> forward tap-major K-major pack [kh*kw*N, C]."""
    n, c, kh, kw = w.shape
    return w.detach().permute(2, 3, 0, 1).reshape(kh * kw * n, c).to(dtype).contiguous()


def pack_conv_weight_dgrad(w, dtype=torch.bfloat16):
    """Data-gradient pack: [kh*kw*C, N] with the taps flipped (dX[p] = sum_t dY[p - off_t] W_t^T)."""
    n, c, kh, kw = w.shape
    return w.detach().flip(2, 3).permute(2, 3, 1, 0).reshape(kh * kw * c, n).to(dtype).contiguous()


def pack_conv_weights(specs):
    """All weight packs of a network in ONE kernel (mv_pack_conv_weights).
    specs: list of (weight [N, C, kh, kw] fp32, Npad, Cpad, want_dgrad) -> list of (fwd [T*Npad, Cpad], dgrad [T*Cpad, Npad] | None),
    the same matrices as pack_conv_weight / pack_conv_weight_dgrad on the zero-padded weight."""
    items = (C.PackItem * len(specs))()
    outs = []
    for it, (w, npad, cpad, want_d) in zip(items, specs):
        w = w.detach()
        assert w.dtype == torch.float32 and w.is_contiguous() and w.is_cuda
        n, c, kh, kw = w.shape
        t = kh * kw
        fwd = torch.empty(t * npad, cpad, device=w.device, dtype=torch.bfloat16)
        dg = torch.empty(t * cpad, npad, device=w.device, dtype=torch.bfloat16) if want_d else None
        it.src, it.dst_fwd, it.dst_dgrad = w.data_ptr(), fwd.data_ptr(), None if dg is None else dg.data_ptr()
        it.N, it.C, it.T, it.Npad, it.Cpad = n, c, t, npad, cpad
        outs.append((fwd, dg))
    C.check(C.lib().mv_pack_conv_weights(items, len(specs), C.stream()), "mv_pack_conv_weights")
    return outs


def tapgemm(A, Wt, T, tap_off, N_total, P, *, Cin=None, BN=None, bias=None, act="none", alpha=1.0, res=None, dact1=None,
            slope1=0.2, out=None, out2=None, out2_pre=False, alpha2=1.0, dact2=None, slope2=0.2, geom=None,
            nchw_out=None, n_valid=0, tag=None, out2_mask=None, dmask2=None, dmask1=None, res_mask=None, res_scale=(1.0, 1.0),
            a2=None, w2=None):
    """out[p, n] = epilogue(sum_t sum_c A[p + tap_off[t], c] * Wt[t*N_total + n, c]); see mv_tapgemm."""
    lib = C.lib()
    Cin = A.shape[1] if Cin is None else Cin
    assert A.dtype == torch.bfloat16 and Wt.dtype == torch.bfloat16 and A.stride(1) == 1 and Wt.is_contiguous()
    assert Wt.shape == (T * N_total, Cin), (Wt.shape, T, N_total, Cin)
    if BN is None:
        BN = 128 if N_total % 128 == 0 else (64 if N_total % 64 == 0 else (32 if N_total % 32 == 0 else 16))
    a = C.TapGemmArgs()
    a.A, a.a_rows, a.a_ld, a.Cin = A.data_ptr(), A.shape[0], A.stride(0), Cin
    a.Wt, a.T = Wt.data_ptr(), T
    for i, o in enumerate(tap_off):
        a.tap_off[i] = int(o)
    a.N_total, a.BN, a.P = N_total, BN, P
    a.bias = None if bias is None else C.ptr(bias)
    a.act, a.alpha = C.ACT[act], float(alpha)
    if nchw_out is not None:
        out = nchw_out
        a.out_mode, a.n_valid, a.out_ld = 1, n_valid, 0
    else:
        if out is None:
            out = torch.empty(P, N_total, device=A.device, dtype=torch.bfloat16)
        a.out_mode, a.out_ld = 0, out.stride(0)
    a.out = out.data_ptr()
    for name, t in (("res", res), ("dact1", dact1), ("dact2", dact2), ("out2", out2)):
        if t is not None:
            assert t.dtype == torch.bfloat16 and t.stride(1) == 1
            setattr(a, name, t.data_ptr())
            setattr(a, name + "_ld", t.stride(0))
    a.slope1, a.slope2, a.out2_pre, a.alpha2 = float(slope1), float(slope2), int(out2_pre), float(alpha2)
    if out2_mask is not None:
        assert out2 is None and out2_mask.dtype == torch.int64 and out2_mask.is_contiguous() and out2_mask.numel() >= mask_rows(P)
        a.out2_mask = out2_mask.data_ptr()
    if dmask1 is not None:
        assert dact1 is None and dmask1.dtype == torch.int64 and dmask1.is_contiguous() and N_total == 64
        a.dmask1 = dmask1.data_ptr()
    if dmask2 is not None:
        assert dact2 is None and dmask2.dtype == torch.int64 and dmask2.is_contiguous() and N_total == 64
        a.dmask2 = dmask2.data_ptr()
    if res_mask is not None:
        assert res is not None and res_mask.dtype == torch.int64 and res_mask.is_contiguous() and N_total == 64
        a.res_mask, a.res_scale_pos, a.res_scale_neg = res_mask.data_ptr(), float(res_scale[0]), float(res_scale[1])
    if a2 is not None:   # fused 1x1 term: out += a2 @ w2^T (plain 64 -> 64 3x3 convolutions only)
        assert a2.dtype == torch.bfloat16 and a2.stride(1) == 1 and w2.dtype == torch.bfloat16 and w2.is_contiguous() and w2.shape == (64, 64)
        a.A2, a.a2_ld, a.W2 = a2.data_ptr(), a2.stride(0), w2.data_ptr()
    if geom is not None:
        a.img_stride, a.Wp, a.W, a.H, a.n_img = geom.S, geom.Wp, geom.W, geom.H, geom.n_img
    kw = {} if tag is None else {"tag": tag}
    C.check(lib.mv_tapgemm(a, C.stream(), **kw), "mv_tapgemm")
    return out


def wgrad(X, G, T, tap_off, P, dW=None, Cin=None, N=None, tag=None, want_db=False, db=None, grad_out=None, n_valid=None):
    """dW[t, n, c] += sum_p G[p, n] * X[p + tap_off[t], c]  (fp32 [T, N, Cin]); see mv_wgrad.
    want_db: also return db[n] = sum_p G[p, n] (bias gradient), fused into the same kernel.
    grad_out: accumulate straight into a gradient tensor in the torch Conv2d layout [n_valid, Cin, kh, kw] (e.g. weight.grad;
    mv_wgrad_nct) instead of a [T, N, Cin] buffer; returns (grad_out, db) / grad_out."""
    import ctypes
    lib = C.lib()
    Cin = X.shape[1] if Cin is None else Cin
    N = G.shape[1] if N is None else N
    assert X.dtype == torch.bfloat16 and G.dtype == torch.bfloat16 and X.stride(1) == 1 and G.stride(1) == 1
    offs = (ctypes.c_int32 * 9)(*[int(o) for o in tap_off])
    kw = {} if tag is None else {"tag": tag}
    if want_db and db is None:
        db = torch.zeros(N, device=X.device, dtype=torch.float32)
    if not want_db:
        db = None
    dbp = None if db is None else db.data_ptr()
    if grad_out is not None:
        n_valid = N if n_valid is None else n_valid
        assert grad_out.dtype == torch.float32 and grad_out.is_contiguous() and grad_out.numel() == n_valid * Cin * T
        for n0 in range(0, N, 128):
            nn = min(128, N - n0)
            Gs = G[:, n0:n0 + nn]
            C.check(lib.mv_wgrad_nct(X.data_ptr(), X.shape[0], X.stride(0), Cin, Gs.data_ptr(), G.shape[0], G.stride(0), nn, T, offs, P,
                                     grad_out.data_ptr(), n0, min(nn, n_valid - n0), dbp, C.stream(), **kw), "mv_wgrad_nct")
        return (grad_out, db) if want_db else grad_out
    if dW is None:
        dW = torch.zeros(T, N, Cin, device=X.device, dtype=torch.float32)
    if N <= 128:
        C.check(lib.mv_wgrad(X.data_ptr(), X.shape[0], X.stride(0), Cin, G.data_ptr(), G.shape[0], G.stride(0), N, T, offs, P,
                             dW.data_ptr(), dbp, C.stream(), **kw), "mv_wgrad")
    else:  # wide layers: 128-column slices of G, each into its rows of dW
        assert N % 128 == 0
        for n0 in range(0, N, 128):
            Gs = G[:, n0:n0 + 128]
            C.check(lib.mv_wgrad_slice(X.data_ptr(), X.shape[0], X.stride(0), Cin, Gs.data_ptr(), G.shape[0], G.stride(0), 128, T,
                                       offs, P, dW.data_ptr(), N, n0, dbp, C.stream(), **kw), "mv_wgrad_slice")
    return (dW, db) if want_db else dW


def unpack_wgrad_add(items):
    """grad[n, c, t] += dW[t, n, c] for every (dW, grad[, swapped]) pair in ONE kernel (mv_unpack_wgrad_add).
    dW: fp32 [T, Npad, Cpad] as produced by `wgrad` (bias gradients: [Npad]); grad: the parameter's contiguous fp32 .grad."""
    arr = (C.UnpackItem * len(items))()
    for it, item in zip(arr, items):
        dW, grad = item[0], item[1]
        swapped = len(item) > 2 and item[2]
        assert dW.dtype == torch.float32 and dW.is_contiguous() and grad.dtype == torch.float32 and grad.is_contiguous()
        if dW.dim() == 1:       # bias
            it.N, it.C, it.T, it.Npad, it.Cpad = grad.numel(), 1, 1, dW.shape[0], 1
        else:
            t = dW.shape[0]
            n, c = grad.shape[0], grad.shape[1]
            it.N, it.C, it.T = n, c, t
            it.Npad, it.Cpad = (dW.shape[2], dW.shape[1]) if swapped else (dW.shape[1], dW.shape[2])
            assert grad.numel() == n * c * t
        it.src, it.dst, it.swapped = dW.data_ptr(), grad.data_ptr(), 1 if swapped else 0
    C.check(C.lib().mv_unpack_wgrad_add(arr, len(items), C.stream()), "mv_unpack_wgrad_add")


class ZeroArena:
    """One zero-filled fp32 buffer handed out in slices: the accumulating outputs (dW, db) of all weight-gradient launches
    of a backward pass share ONE fill kernel instead of one each."""

    def __init__(self, n_floats, device):
        self.buf = torch.zeros(n_floats, device=device, dtype=torch.float32)
        self.off = 0

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= d
        t = self.buf[self.off:self.off + n].view(*shape)
        self.off += (n + 3) // 4 * 4   # keep 16-byte alignment
        assert self.off <= self.buf.numel(), "ZeroArena too small"
        return t


def unpack_conv_wgrad(dW, kh, kw):
    """[kh*kw, N, C] fp32 -> torch Conv2d weight-gradient layout [N, C, kh, kw]."""
    t, n, c = dW.shape
    return dW.view(kh, kw, n, c).permute(2, 3, 0, 1)
