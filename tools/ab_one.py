"""One kernel case at the north-star size, a few launches, for ncu captures: python tools/ab_one.py <lib.so|-> <case> [reps]
cases: headd, wg_b1c0, wg_head, wg_b2c0, c3_c0, c3_c1, c3_c0d, c3_c1d"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multivae_b200 import _cabi  # noqa: E402

if sys.argv[1] not in ("", "-"):
    _cabi.LIB_PATH = os.path.abspath(sys.argv[1])
from multivae_b200.nn import halo as HL  # noqa: E402

case = sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = "cuda"
torch.manual_seed(0)
n_img = 12800


def halo_rand(g, C):
    x = torch.zeros(g.P, C, device=dev, dtype=torch.bfloat16)
    x[: g.n_img * g.S].view(g.n_img, g.H + 1, g.Wp, C)[:, 1:, :g.W] = torch.randn(g.n_img, g.H, g.W, C, device=dev).bfloat16()
    return x


g28, g14, g7 = HL.Geom(n_img, 28, 28), HL.Geom(n_img, 14, 14), HL.Geom(n_img, 7, 7)
if case == "headd":
    gh = halo_rand(g28, 16)
    whd = (torch.randn(9 * 64, 16, device=dev) * 0.05).bfloat16()
    mask = torch.randint(-2 ** 62, 2 ** 62, (HL.mask_rows(g28.P),), device=dev, dtype=torch.int64)
    out = torch.empty(g28.P, 64, device=dev, dtype=torch.bfloat16)
    fn = lambda: HL.tapgemm(gh, whd, 9, g28.taps3x3(), 64, g28.P, alpha=0.1, dmask1=mask, slope1=0.2, geom=g28, out=out)
elif case == "wg_b1c0":
    x7, gg7 = halo_rand(g7, 256), halo_rand(g7, 128)
    dW7 = torch.zeros(9, 128, 256, device=dev)
    fn = lambda: HL.wgrad(x7, gg7, 9, g7.taps3x3(), g7.P, dW=dW7)
elif case == "wg_head":
    o3, gh = halo_rand(g28, 64), halo_rand(g28, 16)
    dWh = torch.zeros(9, 16, 64, device=dev)
    fn = lambda: HL.wgrad(o3, gh, 9, g28.taps3x3(), g28.P, dW=dWh)
elif case == "wg_b2c0":
    x14, gg14 = halo_rand(g14, 128), halo_rand(g14, 64)
    dW14 = torch.zeros(9, 64, 128, device=dev)
    fn = lambda: HL.wgrad(x14, gg14, 9, g14.taps3x3(), g14.P, dW=dW14)
elif case.startswith("c3_"):
    # the four 64 -> 64 3x3 convolutions of the last decoder block (three-taps-per-MMA kernel)
    x = halo_rand(g28, 64)
    r = x.flip(1).contiguous()
    w = (torch.randn(9 * 64, 64, device=dev) * 0.05).bfloat16()
    b = torch.randn(64, device=dev)
    mask = torch.randint(-2 ** 62, 2 ** 62, (HL.mask_rows(g28.P),), device=dev, dtype=torch.int64)
    m2 = torch.empty_like(mask)
    out = torch.empty(g28.P, 64, device=dev, dtype=torch.bfloat16)
    kw = {"c3_c0": dict(bias=b, act="lrelu", out2_mask=m2),
          "c3_c1": dict(bias=b, act="lrelu", alpha=0.1, res=r, out2_mask=m2),
          "c3_c1d": dict(dmask1=mask, slope1=0.2),
          "c3_c0d": dict(res=r, res_mask=mask, res_scale=(10.0, 50.0))}[case]
    fn = lambda: HL.tapgemm(x, w, 9, g28.taps3x3(), 64, g28.P, geom=g28, out=out, **kw)
else:
    raise SystemExit("unknown case " + case)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("done", case)
