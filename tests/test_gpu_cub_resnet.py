"""The 64 x 64 CUB ResNets (multivae_b200/nn/cub.py -> cub_native.py) on the native tensor-core path against the goldens of the
REAL reference modules (tests/golden/nets_{enc,dec}_cub_resnet.pt, models/nn/cub.py:144-293), and against the same modules run
by the library under bf16 autocast on the same GPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("net", ["enc_cub_resnet", "dec_cub_resnet"])
def test_native_cub_resnets_match_reference_goldens(net):
    from tests.test_gpu_gemm import native_vs_library_bf16
    e_nat, e_lib = native_vs_library_bf16(net)
    assert e_nat["out"] <= 3e-2
    keys = ("out", "grad") + (("grad_in", "grad_in_l2") if "grad_in" in e_nat else ())
    for k in keys:
        assert e_nat[k] <= 1.5 * e_lib[k] + 5e-3, (net, k, e_nat[k], e_lib[k])


def test_lrelu_fwd_kernel():
    from multivae_b200 import _cabi as C
    x = (torch.rand(1000, 64, device="cuda") * 2 - 1).bfloat16()
    out = torch.empty_like(x)
    C.check(C.lib().mv_lrelu_fwd(x.data_ptr(), out.data_ptr(), 1000, 64, 0.2, C.stream()), "mv_lrelu_fwd")
    assert torch.equal(out, torch.nn.functional.leaky_relu(x.float(), 0.2).bfloat16())


def test_cub_resnets_train_inside_a_model():
    """One MVTCAE training step with the CUB ResNet pair as the image modality (64 x 64 images next to a 40-attribute vector: the
    SURVEY section 8(d) cfg5 stretch shapes), on the native path; every parameter receives a finite gradient."""
    import multivae_b200 as mb
    from multivae_b200 import nn as N
    torch.manual_seed(0)
    cfg = mb.MVTCAEConfig(n_modalities=2, latent_dim=16, input_dims={"image": (3, 64, 64), "attributes": (40,)})
    enc = {"image": N.CUB_Resnet_Encoder(16), "attributes": N.Encoder_VAE_MLP(mb.BaseAEConfig(input_dim=(40,), latent_dim=16))}
    dec = {"image": N.CUB_Resnet_Decoder(16), "attributes": N.Decoder_AE_MLP(mb.BaseAEConfig(input_dim=(40,), latent_dim=16))}
    m = mb.MVTCAE(cfg, enc, dec).cuda()
    m.compute_dtype = torch.bfloat16
    ds = mb.MultimodalBaseDataset(data={"image": torch.rand(4, 3, 64, 64).cuda(), "attributes": torch.rand(4, 40).cuda()})
    torch.manual_seed(1)
    out = m(ds)
    out.loss.backward()
    assert torch.isfinite(out.loss)
    for k, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    # same step on the library fp32 path: the losses agree to bf16 accuracy
    m.zero_grad()
    m.compute_dtype = torch.float32
    torch.manual_seed(1)
    ref = m(ds).loss
    assert abs(float(out.loss.detach()) - float(ref.detach())) <= 2e-2 * abs(float(ref.detach()))


@pytest.mark.parametrize("lead", [(1,), (2, 3)])
def test_cub_decoder_leading_dimensions_and_no_grad(lead):
    """z of shape [B, L] or [K, B, L] (importance samples): the native decoder flattens the leading dimensions; under no_grad
    (inference API) nothing is saved.  Compared with the library layers in fp32."""
    from multivae_b200 import nn as N
    from multivae_b200.nn import functional as NF
    torch.manual_seed(3)
    dec = N.CUB_Resnet_Decoder(8).cuda()
    z = torch.randn(*lead, 8, device="cuda")
    NF.set_backend("torch")
    try:
        ref = dec(z).reconstruction
    finally:
        NF.set_backend("native")
    try:
        with torch.no_grad():
            got = dec(z).reconstruction
    finally:
        NF.set_backend("auto")
    assert got.shape == ref.shape == (*lead, 3, 64, 64)
    assert float((got.float() - ref).abs().max()) <= 3e-2 * float(ref.abs().max())
