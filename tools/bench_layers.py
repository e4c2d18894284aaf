"""Times the tensor-core layer kernels on the north-star decoder shapes (n_img = M*K*B images per decoder)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multivae_b200.nn import halo as HL

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 12800
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1


def timeit(fn, n=10):
    for _ in range(2):
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
    _a @ _a  # clock ramp before the first measurement
torch.cuda.synchronize()

for ci, (H, cin, cout, T) in enumerate([(28, 64, 64, 9), (14, 128, 64, 9), (14, 64, 64, 9), (7, 256, 128, 9), (7, 128, 128, 9),
                          (14, 128, 64, 1), (7, 256, 128, 1), (28, 64, 16, 9), (28, 16, 64, 9)]):
    if only >= 0 and ci != only:
        continue
    g = HL.Geom(n_img, H, H)
    A = torch.randn(g.P, cin, device="cuda").bfloat16()
    W = (torch.randn(T * cout, cin, device="cuda") * 0.05).bfloat16()
    b = torch.zeros(cout, device="cuda")
    out = torch.empty(g.P, cout, device="cuda", dtype=torch.bfloat16)
    taps = g.taps3x3() if T == 9 else [0]
    ms = timeit(lambda: HL.tapgemm(A, W, T, taps, cout, g.P, bias=b, act="lrelu", out=out, geom=g))
    useful = 2.0 * n_img * H * H * cin * cout * T
    executed = 2.0 * g.P * cin * cout * T
    print(f"H={H:2d} cin={cin:3d} cout={cout:3d} T={T}: {ms:8.3f} ms  useful {useful/ms/1e9:7.1f} TFLOP/s  executed {executed/ms/1e9:7.1f} TFLOP/s  "
          f"in+out {(g.P*(cin+cout)*2)/ms/1e6:7.1f} GB/s")
