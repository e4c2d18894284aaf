"""Pipeline experiments on the 64->64 3x3 tap-GEMM at 28x28 (the dominant shape): which stage bounds each epilogue variant.
  python tools/exp_tapgemm.py [n_img]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multivae_b200.nn import halo as HL

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 12800


def timeit(fn, n=8):
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
    _a @ _a
torch.cuda.synchronize()

H, cin, cout, T = 28, 64, 64, 9
g = HL.Geom(n_img, H, H)
A = torch.randn(g.P, cin, device="cuda").bfloat16()
S = torch.randn(g.P, cout, device="cuda").bfloat16()
W = (torch.randn(T * cout, cin, device="cuda") * 0.05).bfloat16()
b = torch.zeros(cout, device="cuda")
out = torch.empty(g.P, cout, device="cuda", dtype=torch.bfloat16)
out2 = torch.empty(g.P, cout, device="cuda", dtype=torch.bfloat16)
taps = g.taps3x3()
cases = {
    "plain(c0)": lambda: HL.tapgemm(A, W, T, taps, cout, g.P, bias=b, act="lrelu", out=out, geom=g),
    "res+out2(c1)": lambda: HL.tapgemm(A, W, T, taps, cout, g.P, bias=b, act="lrelu", alpha=0.1, res=S, out=out, out2=out2, out2_pre=True, geom=g),
    "dact1(c1d)": lambda: HL.tapgemm(A, W, T, taps, cout, g.P, dact1=S, out=out, geom=g),
    "res(c0d)": lambda: HL.tapgemm(A, W, T, taps, cout, g.P, res=S, out=out, geom=g),
}
envs = [{}, {"MV_TG_IN_STAGES": "2"}, {"MV_TG_IN_STAGES": "1"}, {"MV_TG_DBG": "1"}, {"MV_TG_DBG": "2"}, {"MV_TG_DBG": "3"}, {"MV_TG_DBG": "4"}]
extra = [e for e in sys.argv[2:]]
for e in extra:
    k, v = e.split("=")
    envs.append({k: v})
for name, fn in cases.items():
    for env in envs:
        for k in ("MV_TG_IN_STAGES", "MV_TG_DBG", "MV_TG_X"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ms = timeit(fn)
        print(f"{name:14s} {str(env):32s} {ms:7.3f} ms   {ms*1e-3*1.9e9/ (g.P/128/148):7.0f} cyc/tile@1.9GHz")
