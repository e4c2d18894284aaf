import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multivae_b200.nn import halo as HL


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

_a = torch.randn(8192, 8192, device="cuda").bfloat16()
for _ in range(60):
    _a @ _a
torch.cuda.synchronize()
for (H, n_img) in [(28, 800), (28, 3200), (28, 6400), (28, 12800), (14, 12800), (14, 25600), (14, 47800), (7, 12800), (7, 170000)]:
    cin = cout = 64
    g = HL.Geom(n_img, H, H)
    A = torch.randn(g.P, cin, device="cuda").bfloat16()
    W = (torch.randn(9 * cout, cin, device="cuda") * 0.05).bfloat16()
    b = torch.zeros(cout, device="cuda")
    out = torch.empty(g.P, cout, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: HL.tapgemm(A, W, 9, g.taps3x3(), cout, g.P, bias=b, act="lrelu", out=out, geom=g))
    tiles = (g.P + 127) // 128
    print(f"H={H:2d} n_img={n_img:6d} rows={g.P:9d} tiles/CTA={tiles/148:7.1f}: {ms:7.3f} ms  {ms*1e3/(tiles/148):6.3f} us/tile  executed {2.0*g.P*cin*cout*9/ms/1e9:6.1f} TFLOP/s")
