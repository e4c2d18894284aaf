"""tcgen05 weight-gradient kernel (mv_wgrad) against autograd of a plain PyTorch fp32 convolution on the
same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.rand(*shape, device="cuda", generator=g) * 2 - 1) * scale


@pytest.mark.parametrize("n_img,H,cin,cout,k", [(37, 28, 64, 64, 3), (300, 28, 64, 64, 3), (23, 14, 128, 64, 3),
                                                (41, 7, 256, 128, 3), (19, 7, 128, 128, 3), (29, 14, 128, 64, 1),
                                                (31, 7, 256, 128, 1), (33, 28, 64, 16, 3),
                                                # CUB ResNet stages (64-pixel rows: two-box windows)
                                                (4, 64, 64, 64, 3), (3, 64, 64, 16, 3), (3, 32, 128, 64, 3), (3, 32, 64, 128, 3),
                                                (3, 16, 256, 128, 3), (2, 16, 128, 256, 3), (3, 32, 64, 128, 1), (3, 16, 128, 256, 1)])
def test_conv_weight_gradient(n_img, H, cin, cout, k):
    from multivae_b200.nn import halo as HL
    x = _rnd(n_img, cin, H, H, seed=1).bfloat16()
    gy = _rnd(n_img, cout, H, H, seed=2).bfloat16()
    X, g = HL.to_halo(x)
    G, _ = HL.to_halo(gy)
    taps = g.taps3x3() if k == 3 else [0]
    dW, db = HL.wgrad(X, G, k * k, taps, g.P, want_db=True)
    ref_db = gy.float().sum((0, 2, 3))
    assert float((db - ref_db).abs().max()) / float(ref_db.abs().max()) < 2e-3
    w = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
    F.conv2d(x.float(), w, padding=k // 2).backward(gy.float())
    got = HL.unpack_conv_wgrad(dW, k, k)
    err = float((got - w.grad).abs().max()) / float(w.grad.abs().max())
    assert err < 2e-3, err


def test_accumulates_into_existing_gradient():
    from multivae_b200.nn import halo as HL
    x = _rnd(500, 64, seed=3).bfloat16()
    gy = _rnd(500, 64, seed=4).bfloat16()
    base = _rnd(1, 64, 64, seed=5)
    dW = HL.wgrad(x, gy, 1, [0], 500, dW=base.clone())
    ref = base[0] + gy.float().t() @ x.float()
    assert float((dW[0] - ref).abs().max()) / float(ref.abs().max()) < 2e-3
