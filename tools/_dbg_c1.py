import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from multivae_b200.nn import halo as HL
def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.rand(*shape, device="cuda", generator=g) * 2 - 1) * scale
n_img,H,cin,cout=37,28,64,64
x = _rnd(n_img, cin, H, H, seed=1).bfloat16()
w = _rnd(cout, cin, 3, 3, seed=2, scale=cin ** -0.5).bfloat16()
b = _rnd(cout, seed=3)
r = _rnd(n_img, cout, H, H, seed=4).bfloat16()
A, g = HL.to_halo(x)
R, _ = HL.to_halo(r)
y = F.leaky_relu(F.conv2d(x.float(), w.float(), b, padding=1), 0.2)
Y,_ = HL.to_halo(y, dtype=torch.float32)
O,_ = HL.to_halo(r.float() + 0.1*y, dtype=torch.float32)
for env in ({}, {"MV_NO_CONV3":"1"}):
    os.environ.pop("MV_NO_CONV3", None); os.environ.update(env)
    act = torch.zeros(g.P, cout, device="cuda", dtype=torch.bfloat16)
    out = HL.tapgemm(A, HL.pack_conv_weight(w), 9, g.taps3x3(), cout, g.P, bias=b, act="lrelu", alpha=0.1, res=R, out2=act, out2_pre=True, geom=g)
    torch.cuda.synchronize()
    for name, got, ref in (("act", act, Y), ("out", out, O)):
        err = (got.float()-ref).abs()
        bad = (err > 0.05).nonzero()
        print(env, name, "max err", float(err.max()), "n bad", bad.shape[0], "of", err.numel())
        if bad.shape[0]:
            rows = bad[:,0].unique()
            print("  bad rows (first 40):", rows[:40].tolist(), " rows mod 126:", sorted(set((rows % 126).tolist()))[:40])
            print("  bad cols:", bad[:,1].unique().tolist()[:70])
