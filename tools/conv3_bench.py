"""Times the four 64 -> 64 3x3 convolution variants of the last decoder block at the north-star size (12800 images, 28x28) with CUDA
events: python tools/conv3_bench.py [reps] [path/to/lib.so]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multivae_b200 import _cabi  # noqa: E402

if len(sys.argv) > 2:
    _cabi.LIB_PATH = os.path.abspath(sys.argv[2])
from multivae_b200.nn import halo as HL  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_img, H = 12800, 28
g = HL.Geom(n_img, H, H)
dev = "cuda"
torch.manual_seed(0)
x = torch.zeros(g.P, 64, device=dev, dtype=torch.bfloat16)
x[: n_img * g.S].view(n_img, H + 1, g.Wp, 64)[:, 1:, :H] = torch.randn(n_img, H, H, 64, device=dev).bfloat16()
r = x.flip(1).contiguous()
w = (torch.randn(9 * 64, 64, device=dev) * 0.05).bfloat16()
b = torch.randn(64, device=dev)
mask = torch.randint(-2 ** 62, 2 ** 62, (HL.mask_rows(g.P),), device=dev, dtype=torch.int64)
m2 = torch.empty_like(mask)
taps = g.taps3x3()
out = torch.empty(g.P, 64, device=dev, dtype=torch.bfloat16)
variants = {
    "c0  (bias, lrelu, mask2)": dict(bias=b, act="lrelu", out2_mask=m2),
    "c1  (bias, lrelu, res, mask2)": dict(bias=b, act="lrelu", alpha=0.1, res=r, out2_mask=m2),
    "c1d (dmask1)": dict(dmask1=mask, slope1=0.2),
    "c0d (res, res_mask)": dict(res=r, res_mask=mask, res_scale=(10.0, 50.0)),
    "plain": dict(),
    "res (b2.c0d halves)": dict(res=r),
    "bias, lrelu (encoder c0)": dict(bias=b, act="lrelu"),
    "bias, lrelu, res, out2": dict(bias=b, act="lrelu", alpha=0.1, res=r, out2=torch.empty(g.P, 64, device=dev, dtype=torch.bfloat16), out2_pre=True),
    "dact1": dict(dact1=r, slope1=0.2),
}
flops = 2.0 * n_img * H * H * 64 * 64 * 9
print("library:", _cabi.LIB_PATH)
for name, kw in variants.items():
    for _ in range(3):
        HL.tapgemm(x, w, 9, taps, 64, g.P, geom=g, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        HL.tapgemm(x, w, 9, taps, 64, g.P, geom=g, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:32s} {us:8.1f} us   {flops / us / 1e6:7.1f} TFLOP/s useful   checksum {out.float().abs().sum().item():.6e}", flush=True)
