"""A/B timing of single kernels at the north-star size (12800 decoder images) under a chosen build of the library:
    python tools/ab_bench.py [path/to/lib.so] [reps]
Prints CUDA-event times per launch and a checksum of every result, so that two builds can be compared line by line
(the weight-gradient kernels reduce with fp32 atomics: their checksums agree to rounding, not bit for bit)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multivae_b200 import _cabi  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] not in ("", "-"):
    _cabi.LIB_PATH = os.path.abspath(sys.argv[1])
from multivae_b200.nn import halo as HL  # noqa: E402

reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda"
torch.manual_seed(0)
n_img = 12800


def halo_rand(g, C, scale=1.0):
    x = torch.zeros(g.P, C, device=dev, dtype=torch.bfloat16)
    x[: g.n_img * g.S].view(g.n_img, g.H + 1, g.Wp, C)[:, 1:, :g.W] = (torch.randn(g.n_img, g.H, g.W, C, device=dev) * scale).bfloat16()
    return x


def timed(name, fn, work=None, unit="TFLOP/s"):
    for _ in range(2):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    rr = r[0] if isinstance(r, tuple) else r
    chk = rr.float().abs().sum().item()
    extra = "" if work is None else f"  {work / us / 1e6:8.1f} {unit}"
    print(f"{name:34s} {us:9.1f} us{extra}   checksum {chk:.6e}", flush=True)


print("library:", _cabi.LIB_PATH, flush=True)
g28, g14, g7 = HL.Geom(n_img, 28, 28), HL.Geom(n_img, 14, 14), HL.Geom(n_img, 7, 7)

# ---- head.d: data gradient of the image head (16 -> 64 at 28x28, sign-mask epilogue) ----
gh = halo_rand(g28, 16)
whd = (torch.randn(9 * 64, 16, device=dev) * 0.05).bfloat16()
mask = torch.randint(-2 ** 62, 2 ** 62, (HL.mask_rows(g28.P),), device=dev, dtype=torch.int64)
out = torch.empty(g28.P, 64, device=dev, dtype=torch.bfloat16)
timed("head.d (16->64, dmask1)", lambda: HL.tapgemm(gh, whd, 9, g28.taps3x3(), 64, g28.P, alpha=0.1, dmask1=mask, slope1=0.2, geom=g28, out=out),
      2.0 * n_img * 784 * 16 * 64 * 9)
bias = torch.randn(64, device=dev)
timed("e.img-like (16->64, bias)", lambda: HL.tapgemm(gh, whd, 9, g28.taps3x3(), 64, g28.P, bias=bias, geom=g28, out=out),
      2.0 * n_img * 784 * 16 * 64 * 9)

# ---- weight gradient of the image head (X 64 ch, G 16 columns) ----
o3 = halo_rand(g28, 64)
dWh = torch.zeros(9, 16, 64, device=dev)
dbh = torch.zeros(16, device=dev)


def wg_head():
    dWh.zero_(); dbh.zero_()
    return HL.wgrad(o3, gh, 9, g28.taps3x3(), g28.P, want_db=True, dW=dWh, db=dbh)


timed("wgrad head (64 x 16 @28)", wg_head, 2.0 * n_img * 784 * 16 * 64 * 9)
gw = torch.zeros(3, 64, 3, 3, device=dev)


def wg_head_nct():
    gw.zero_(); dbh.zero_()
    return HL.wgrad(o3, gh, 9, g28.taps3x3(), g28.P, want_db=True, grad_out=gw, n_valid=3, db=dbh)


timed("wgrad head nct (n_valid 3)", wg_head_nct)
del o3, gh, out

# ---- weight gradient 256 -> 128 at 7x7 (18 accumulators, 5 passes) ----
x7 = halo_rand(g7, 256)
gg7 = halo_rand(g7, 128)
dW7 = torch.zeros(9, 128, 256, device=dev)
db7 = torch.zeros(128, device=dev)


def wg_b1c0():
    dW7.zero_(); db7.zero_()
    return HL.wgrad(x7, gg7, 9, g7.taps3x3(), g7.P, want_db=True, dW=dW7, db=db7)


timed("wgrad b1.c0 (256 x 128 @7)", wg_b1c0, 2.0 * n_img * 49 * 256 * 128 * 9)
# reference value of a few entries (fp32 torch on a slice of the images would be slow: compare builds instead)
print("  dW7[4, 5, 7], dW7[0, 100, 200], db7[3]:", dW7[4, 5, 7].item(), dW7[0, 100, 200].item(), db7[3].item())
x7s = halo_rand(g7, 128)
dW7b = torch.zeros(9, 128, 128, device=dev)


def wg_b1c1():
    dW7b.zero_()
    return HL.wgrad(x7s, gg7, 9, g7.taps3x3(), g7.P, dW=dW7b)


timed("wgrad b1.c1 (128 x 128 @7)", wg_b1c1, 2.0 * n_img * 49 * 128 * 128 * 9)
del x7, gg7, x7s

# ---- weight gradient 128 -> 64 at 14x14 and 64 -> 64 at 28x28 (unchanged kernels: drift check between the runs) ----
x14 = halo_rand(g14, 128)
gg14 = halo_rand(g14, 64)
dW14 = torch.zeros(9, 64, 128, device=dev)


def wg_b2c0():
    dW14.zero_()
    return HL.wgrad(x14, gg14, 9, g14.taps3x3(), g14.P, dW=dW14)


timed("wgrad b2.c0 (128 x 64 @14)", wg_b2c0, 2.0 * n_img * 196 * 128 * 64 * 9)

# ---- decoder fc: [12800, 64] x [16384, 64]^T + bias -> bf16 [12800, 16384] (K = 64: one K block per tile) ----
del x14, gg14
from multivae_b200.nn import linear_native as LN  # noqa: E402

zin = torch.randn(n_img, 64, device=dev).bfloat16()
wfc = (torch.randn(16384, 64, device=dev) * 0.1).bfloat16()
bfc = torch.randn(16384, device=dev)
ofc = torch.empty(n_img, 16384, device=dev, dtype=torch.bfloat16)
timed("dec.fc (12800 x 16384 x 64)", lambda: LN.gemm(zin, wfc, n_img, 16384, 64, ofc, bias=bfc), 2.0 * n_img * 16384 * 64)
ref = (zin[:64].float() @ wfc.float().t() + bfc)
print("  dec.fc max |err| on the first 64 rows:", (ofc[:64].float() - ref).abs().max().item(), "of", ref.abs().max().item())
ofr = torch.empty(n_img, 16384, device=dev, dtype=torch.bfloat16)
timed("dec.fc + relu", lambda: LN.gemm(zin, wfc, n_img, 16384, 64, ofr, bias=bfc, act="relu"), 2.0 * n_img * 16384 * 64)
print("  relu max |err|:", (ofr[:64].float() - ref.clamp_min(0)).abs().max().item())
