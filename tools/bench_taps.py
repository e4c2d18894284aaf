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
P = 5_000_000
cin = cout = 64
A = torch.randn(P, cin, device="cuda").bfloat16()
W = (torch.randn(9 * cout, cin, device="cuda") * 0.05).bfloat16()
out = torch.empty(P, cout, device="cuda", dtype=torch.bfloat16)
g28 = HL.Geom(1, 28, 28)
g14 = HL.Geom(1, 14, 14)
only = sys.argv[1].split(",") if len(sys.argv) > 1 else None
cases = {
    "wp29": g28.taps3x3(), "wp15": g14.taps3x3(), "mult8": [-32, -24, -16, -8, 0, 8, 16, 24, 32], "zeros": [0] * 9,
    "wp29_lo_only": [-30, -29, -28, -1, 0, 0, 0, 0, 0], "span60_aligned": [-32, -24, -16, 0, 0, 0, 16, 24, 32 - 4],
    "wp31": [(r - 1) * 31 + (s - 1) for r in range(3) for s in range(3)], "wp33": [(r - 1) * 33 + (s - 1) for r in range(3) for s in range(3)],
    "wp23": [(r - 1) * 23 + (s - 1) for r in range(3) for s in range(3)],
}
for name, taps in cases.items():
    if only and name not in only:
        continue
    ms = timeit(lambda: HL.tapgemm(A, W, 9, taps, cout, P, out=out))
    tiles = (P + 127) // 128
    print(f"{name:16s} R={128 - min(taps) + max(taps):4d}: {ms:7.3f} ms  {ms*1e3/(tiles/148):6.3f} us/tile")
