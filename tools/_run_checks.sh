timeout 600 python -m pytest tests/test_gpu_wgrad.py tests/test_gpu_resnet_decoder.py -x -q 2>&1 | tail -5
cat > /tmp/wg_time.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from multivae_b200.nn import halo as HL
def timeit(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = HL.Geom(12800, 28, 28)
X = torch.randn(g.P, 64, device="cuda").bfloat16(); G = torch.randn(g.P, 16, device="cuda").bfloat16()
dW = torch.zeros(9, 16, 64, device="cuda")
for env in ({}, {"MV_WG_NO_PAIR": "1"}):
    os.environ.pop("MV_WG_NO_PAIR", None); os.environ.update(env)
    print(env, timeit(lambda: HL.wgrad(X, G, 9, g.taps3x3(), g.P, dW=dW, want_db=True)), "ms")
PY
timeout 100 python /tmp/wg_time.py
