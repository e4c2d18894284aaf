"""Native (tcgen05) ResNet encoder against the plain PyTorch fp32 modules with the same weights (and against torch's own
bf16 autocast as the yardstick for bf16-operand error): outputs and every parameter gradient."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm()) / (float(b.norm()) + 1e-12)


def _run(enc, x, gy, mode):
    from multivae_b200.nn import functional as NF
    for p in enc.parameters():
        p.grad = None
    NF.set_backend("native" if mode == "native" else "torch")
    try:
        if mode == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                o = enc(x)
        else:
            o = enc(x)
        outs = [o.embedding, o.log_covariance, o.style_embedding, o.style_log_covariance]
        sum((t.float() * g).sum() for t, g in zip(outs, gy)).backward()
    finally:
        NF.set_backend("auto")
    return [t.detach().float() for t in outs], {k: p.grad.clone() for k, p in enc.named_parameters()}


@pytest.mark.parametrize("n_img", [5, 64])
def test_encoder_matches_torch_fp32(n_img):
    from multivae_b200.nn import EncoderResnetMMNIST
    torch.manual_seed(0)
    enc = EncoderResnetMMNIST(32, 32).cuda()
    x = torch.rand(n_img, 3, 28, 28, device="cuda")
    gy = [torch.randn(n_img, 32, device="cuda") for _ in range(4)]
    o_ref, g_ref = _run(enc, x, gy, "fp32")
    o_lib, g_lib = _run(enc, x, gy, "bf16")
    o_nat, g_nat = _run(enc, x, gy, "native")
    for i in range(4):
        e, l = _rel(o_nat[i], o_ref[i]), _rel(o_lib[i], o_ref[i])
        print(f"out{i:<37d} native {e:.4f}   torch-bf16 {l:.4f}")
        assert e < max(1.5e-2, 2.5 * l), (i, e, l)
    for k in g_ref:
        e, l = _rel(g_nat[k], g_ref[k]), _rel(g_lib[k], g_ref[k])
        print(f"{k:40s} native {e:.4f}   torch-bf16 {l:.4f}")
        assert e < max(3e-2, 2.5 * l), (k, e, l)


@pytest.mark.parametrize("mode", ["1", "nct"])
def test_encoder_in_place_gradient_handover_equals_autograd_path(mode):
    """Pre-allocated .grad tensors: the batched mv_unpack_wgrad_add hand-over ("1") / in-kernel accumulation ("nct") give the
    same gradients as returning them to autograd ("0"), on top of an existing gradient value."""
    import os
    from multivae_b200.nn import EncoderResnetMMNIST
    from multivae_b200.nn import functional as NF
    torch.manual_seed(2)
    enc = EncoderResnetMMNIST(32, 32).cuda()
    x = torch.rand(37, 3, 28, 28, device="cuda")
    gy = [torch.randn(37, 32, device="cuda") for _ in range(4)]

    def run():
        o = enc(x)
        outs = [o.embedding, o.log_covariance, o.style_embedding, o.style_log_covariance]
        sum((t.float() * g).sum() for t, g in zip(outs, gy)).backward()

    NF.set_backend("native")
    try:
        os.environ["MULTIVAE_B200_DIRECT_GRADS"] = "0"
        for p in enc.parameters():
            p.grad = None
        run()
        ref = {k: p.grad.clone() for k, p in enc.named_parameters()}
        os.environ["MULTIVAE_B200_DIRECT_GRADS"] = mode
        for p in enc.parameters():
            p.grad = torch.full_like(p, 0.25)
        run()
    finally:
        os.environ.pop("MULTIVAE_B200_DIRECT_GRADS", None)
        NF.set_backend("auto")
    for k, p in enc.named_parameters():
        err = float((p.grad - 0.25 - ref[k]).abs().max()) / (float(ref[k].abs().max()) + 1e-12)
        assert err < 2e-3, (k, err)
