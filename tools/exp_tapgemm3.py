"""Quick A/B of MV_TG_DBG switches on the 64->64 3x3 tap-GEMM at 28x28."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multivae_b200.nn import halo as HL
n_img = 12800
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
g = HL.Geom(n_img, 28, 28)
A = torch.randn(g.P, 64, device="cuda").bfloat16()
S = torch.randn(g.P, 64, device="cuda").bfloat16()
W = (torch.randn(9 * 64, 64, device="cuda") * 0.05).bfloat16()
b = torch.zeros(64, device="cuda")
out = torch.empty(g.P, 64, device="cuda", dtype=torch.bfloat16)
out2 = torch.empty(g.P, 64, device="cuda", dtype=torch.bfloat16)
taps = g.taps3x3()
cases = {
    "plain(c0)": lambda: HL.tapgemm(A, W, 9, taps, 64, g.P, bias=b, act="lrelu", out=out, geom=g),
    "res+out2(c1)": lambda: HL.tapgemm(A, W, 9, taps, 64, g.P, bias=b, act="lrelu", alpha=0.1, res=S, out=out, out2=out2, out2_pre=True, geom=g),
    "dact1(c1d)": lambda: HL.tapgemm(A, W, 9, taps, 64, g.P, dact1=S, out=out, geom=g),
}
envs = [dict(kv.split("=") for kv in a.split(",")) if a != "-" else {} for a in sys.argv[1:]] or [{}]
for name, fn in cases.items():
    for env in envs:
        for k in list(os.environ):
            if k.startswith("MV_TG_") or k == "MV_NO_CONV3":
                os.environ.pop(k)
        os.environ.update(env)
        if "MV_TG_DBG" not in os.environ:
            os.environ["MV_TG_DBG"] = "32"   # no-op bit: only enables the cycle counter
        ms = timeit(fn)
        import ctypes
        from multivae_b200 import _cabi
        cyc, tiles = ctypes.c_longlong(0), ctypes.c_longlong(0)
        getattr(_cabi.lib()._raw, "mv_debug_tg_clk" if "MV_NO_CONV3" in os.environ else "mv_debug_c3_clk")(ctypes.byref(cyc), ctypes.byref(tiles))
        ct = cyc.value / max(1, tiles.value)
        print(f"{name:14s} {str(env):40s} {ms:7.3f} ms   {ct:7.0f} true cyc/tile  -> clock {cyc.value/ (ms*1e-3)/1e9:5.2f} GHz", flush=True)
