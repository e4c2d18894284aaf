"""One native ResNet decoder forward + backward on the north-star image count (M*K*B = 12800 images) for ncu:
  ncu --set full -k regex:mv:: ... python tools/ncu_decoder.py [n_img] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multivae_b200.nn import DecoderResnetMMNIST
from multivae_b200.nn import functional as NF

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 12800
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(0)
dec = DecoderResnetMMNIST(64).cuda()
NF.set_backend("native")
z = torch.randn(n_img, 64, device="cuda", requires_grad=True)
gy = (torch.randn(n_img, 3, 28, 28, device="cuda") * 0.1).to(torch.bfloat16)
for _ in range(reps):
    r = dec(z).reconstruction
    r.backward(gy)
torch.cuda.synchronize()
print("done")
