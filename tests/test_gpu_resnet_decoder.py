"""Native (tcgen05) ResNet decoder stack against the plain PyTorch fp32 modules with the same weights:
reconstruction, gradient w.r.t. the latent input and every parameter gradient (bf16-operand tolerance)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    """relative L2 error (bf16 operands, fp32 accumulation: ~1e-2 through the 11-layer stack)"""
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm()) / (float(b.norm()) + 1e-12)


@pytest.mark.parametrize("n_img", [7, 96])
def test_decoder_forward_backward_matches_torch_fp32(n_img):
    from multivae_b200.nn import DecoderResnetMMNIST
    from multivae_b200.nn import functional as NF
    torch.manual_seed(0)
    dec = DecoderResnetMMNIST(64).cuda()
    z = torch.randn(n_img, 64, device="cuda")
    gy = torch.randn(n_img, 3, 28, 28, device="cuda") * 0.1

    NF.set_backend("torch")
    z0 = z.clone().requires_grad_(True)
    r0 = dec(z0).reconstruction
    r0.backward(gy)
    ref = {k: p.grad.clone() for k, p in dec.named_parameters()}
    for p in dec.parameters():
        p.grad = None

    # the same modules under torch bf16 autocast (cuDNN/cuBLAS): the yardstick for bf16-operand error
    with torch.autocast("cuda", dtype=torch.bfloat16):
        z2 = z.clone().requires_grad_(True)
        r2 = dec(z2).reconstruction
    r2.backward(gy.to(r2.dtype))
    lib = {k: p.grad.clone() for k, p in dec.named_parameters()}
    lib_err = {"recon": _rel(r2, r0), "z.grad": _rel(z2.grad, z0.grad)}
    lib_err.update({k: _rel(lib[k], ref[k]) for k in ref})
    for p in dec.parameters():
        p.grad = None

    NF.set_backend("native")
    try:
        z1 = z.clone().requires_grad_(True)
        r1 = dec(z1).reconstruction
        assert r1.dtype == torch.bfloat16 and r1.shape == r0.shape
        r1.backward(gy.to(torch.bfloat16))
    finally:
        NF.set_backend("auto")
    errs = {"recon": _rel(r1, r0), "z.grad": _rel(z1.grad, z0.grad)}
    for k, p in dec.named_parameters():
        assert p.grad is not None, k
        errs[k] = _rel(p.grad, ref[k])
    for k in errs:
        print(f"{k:40s} native {errs[k]:.4f}   torch-bf16 {lib_err[k]:.4f}")
    assert errs["recon"] < 1e-2, errs["recon"]
    for k, e in errs.items():
        assert e < max(3e-2, 2.5 * lib_err[k]), (k, e, lib_err[k])


@pytest.mark.parametrize("mode", ["1", "nct"])
def test_decoder_direct_gradient_accumulation_equals_autograd_path(mode):
    """With pre-allocated fp32 .grad tensors (the trainer's flat buffer) the gradients are added in place — by one
    mv_unpack_wgrad_add launch ("1", the default) or by the weight-gradient kernels themselves ("nct") — and the autograd
    Function returns None: same gradients as the path that returns them to autograd ("0")."""
    import os
    from multivae_b200.nn import DecoderResnetMMNIST
    from multivae_b200.nn import functional as NF
    torch.manual_seed(1)
    dec = DecoderResnetMMNIST(64).cuda()
    z = torch.randn(33, 64, device="cuda")
    gy = (torch.randn(33, 3, 28, 28, device="cuda") * 0.1).to(torch.bfloat16)
    NF.set_backend("native")
    try:
        os.environ["MULTIVAE_B200_DIRECT_GRADS"] = "0"
        dec(z).reconstruction.backward(gy)
        ref = {k: p.grad.clone() for k, p in dec.named_parameters()}
        os.environ["MULTIVAE_B200_DIRECT_GRADS"] = mode
        for p in dec.parameters():
            p.grad = torch.full_like(p, 0.5)       # accumulation on top of an existing gradient
        dec(z).reconstruction.backward(gy)
    finally:
        os.environ.pop("MULTIVAE_B200_DIRECT_GRADS", None)
        NF.set_backend("auto")
    for k, p in dec.named_parameters():
        got = p.grad - 0.5
        err = float((got - ref[k]).abs().max()) / (float(ref[k].abs().max()) + 1e-12)
        assert err < 2e-3, (k, err)
