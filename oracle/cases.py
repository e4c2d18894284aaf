"""Golden-case definitions shared by oracle/make_golden.py (runs the REAL reference, build container
only) and tests/ (replays the same cases through the oracle port and the CUDA path).
TEST INFRASTRUCTURE ONLY."""
import torch

# name -> spec.  Networks: the reference's default MLP architectures (what Model(config) builds when encoders/decoders are
# None) unless the spec carries `arch` = {modality: "mlp" | "svhn" | "conv_mmnist" | "resnet_mmnist"}; weights from
# oracle.port.nets.synth_state_dict(seed).
_DIMS3 = {"m0": (3, 8, 8), "m1": (1, 6, 6), "m2": (10,)}
_DIST3 = {"m0": "laplace", "m1": "normal", "m2": "normal"}
_PAR3 = {"m0": {"scale": 0.75}, "m1": {"scale": 0.5}, "m2": {}}

CASES = {
    "mmvaeplus_dreg": dict(model="mmvaeplus", dims=_DIMS3, B=6, cfg=dict(K=4, latent_dim=5, modalities_specific_dim=3, beta=2.5, loss="dreg_looser", prior_and_posterior_dist="laplace_with_softmax", decoders_dist=_DIST3, decoder_dist_params=_PAR3, uses_likelihood_rescaling=True)),
    "mmvaeplus_iwae": dict(model="mmvaeplus", dims=_DIMS3, B=6, cfg=dict(K=3, latent_dim=5, modalities_specific_dim=3, beta=1.0, loss="iwae_looser", prior_and_posterior_dist="laplace_with_softmax", decoders_dist=_DIST3, decoder_dist_params=_PAR3, learn_shared_prior=True)),
    "mmvaeplus_normal": dict(model="mmvaeplus", dims=_DIMS3, B=5, cfg=dict(K=2, latent_dim=4, modalities_specific_dim=4, beta=1.5, loss="dreg_looser", prior_and_posterior_dist="normal", decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mmvaeplus_softplus": dict(model="mmvaeplus", dims=_DIMS3, B=5, cfg=dict(K=2, latent_dim=4, modalities_specific_dim=4, beta=1.0, loss="iwae_looser", prior_and_posterior_dist="normal_with_softplus", decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mmvaeplus_masked": dict(model="mmvaeplus", dims=_DIMS3, B=6, masks=True, cfg=dict(K=4, latent_dim=5, modalities_specific_dim=3, beta=2.5, loss="dreg_looser", prior_and_posterior_dist="laplace_with_softmax", decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mmvae_dreg": dict(model="mmvae", dims=_DIMS3, B=6, cfg=dict(K=4, latent_dim=5, loss="dreg_looser", prior_and_posterior_dist="laplace_with_softmax", decoders_dist=_DIST3, decoder_dist_params=_PAR3, uses_likelihood_rescaling=True)),
    "mmvae_iwae": dict(model="mmvae", dims=_DIMS3, B=6, cfg=dict(K=3, latent_dim=5, loss="iwae_looser", prior_and_posterior_dist="normal", decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mmvae_masked": dict(model="mmvae", dims=_DIMS3, B=6, masks=True, cfg=dict(K=3, latent_dim=5, loss="iwae_looser", prior_and_posterior_dist="laplace_with_softmax", decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mvtcae": dict(model="mvtcae", dims=_DIMS3, B=6, cfg=dict(latent_dim=5, alpha=0.1, beta=2.5, decoders_dist=_DIST3, decoder_dist_params=_PAR3, uses_likelihood_rescaling=True)),
    "mvtcae_masked": dict(model="mvtcae", dims=_DIMS3, B=6, masks=True, cfg=dict(latent_dim=5, alpha=0.3, beta=1.0, decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mvae": dict(model="mvae", dims=_DIMS3, B=6, fwd=dict(epoch=12, batch_ratio=0.5), cfg=dict(latent_dim=5, k=0, beta=2.0, warmup=10, decoders_dist=_DIST3, decoder_dist_params=_PAR3, uses_likelihood_rescaling=True)),
    "mvae_warmup_k1": dict(model="mvae", dims=_DIMS3, B=6, fwd=dict(epoch=3, batch_ratio=0.25), np_seed=7, cfg=dict(latent_dim=5, k=1, beta=1.0, warmup=10, decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mvae_masked": dict(model="mvae", dims=_DIMS3, B=6, masks=True, fwd=dict(epoch=12, batch_ratio=0.5), cfg=dict(latent_dim=5, k=0, beta=1.0, warmup=10, decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mopoe": dict(model="mopoe", dims=_DIMS3, B=9, cfg=dict(latent_dim=5, beta=2.5, decoders_dist=_DIST3, decoder_dist_params=_PAR3, uses_likelihood_rescaling=True)),
    "mopoe_5mod": dict(model="mopoe", dims={f"m{i}": (2, 4, 4) for i in range(5)}, B=40, cfg=dict(latent_dim=6, beta=2.5, decoders_dist={f"m{i}": "laplace" for i in range(5)}, decoder_dist_params={f"m{i}": {"scale": 0.75} for i in range(5)})),
    "mopoe_masked": dict(model="mopoe", dims=_DIMS3, B=9, masks=True, cfg=dict(latent_dim=5, beta=1.0, decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    # modality-specific latent spaces in MoPoE (mopoe_model.py:57-74,171-225) and a categorical decoder (base_utils.py:28-59)
    "mopoe_private": dict(model="mopoe", dims=_DIMS3, B=9, cfg=dict(latent_dim=5, beta=2.5, beta_style=2.0, modalities_specific_dim={"m0": 3, "m1": 2, "m2": 4}, decoders_dist=_DIST3, decoder_dist_params=_PAR3, uses_likelihood_rescaling=True)),
    "mopoe_private_masked": dict(model="mopoe", dims=_DIMS3, B=9, masks=True, cfg=dict(latent_dim=5, beta=1.0, beta_style=0.5, modalities_specific_dim={"m0": 3, "m1": 2, "m2": 4}, decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "mvtcae_categorical": dict(model="mvtcae", dims={"m0": (3, 8, 8), "m1": (1, 6, 6), "m2": (4, 5)}, B=6, onehot=["m2"], cfg=dict(latent_dim=5, alpha=0.1, beta=2.5, decoders_dist={"m0": "laplace", "m1": "bernoulli", "m2": "categorical"}, decoder_dist_params={"m0": {"scale": 0.75}, "m1": {}, "m2": {}}, uses_likelihood_rescaling=True)),
    "mmvaeplus_categorical": dict(model="mmvaeplus", dims={"m0": (3, 8, 8), "m2": (4, 5)}, B=5, onehot=["m2"], cfg=dict(K=3, latent_dim=5, modalities_specific_dim=3, beta=2.5, loss="dreg_looser", prior_and_posterior_dist="laplace_with_softmax", decoders_dist={"m0": "laplace", "m2": "categorical"}, decoder_dist_params={"m0": {"scale": 0.75}, "m2": {}})),
    # CMVAE (MMVAE+ with a mixture-of-clusters prior) and CRMVAE (MVTCAE aggregation + unimodal reconstruction terms): SURVEY 8(f) rank 3
    "cmvae_dreg": dict(model="cmvae", dims=_DIMS3, B=6, cfg=dict(K=4, latent_dim=5, modalities_specific_dim=3, beta=2.5, loss="dreg_looser", number_of_clusters=4, prior_and_posterior_dist="laplace_with_softmax", decoders_dist=_DIST3, decoder_dist_params=_PAR3, uses_likelihood_rescaling=True)),
    "cmvae_iwae_normal": dict(model="cmvae", dims=_DIMS3, B=5, cfg=dict(K=3, latent_dim=4, modalities_specific_dim=4, beta=1.0, loss="iwae_looser", number_of_clusters=3, prior_and_posterior_dist="normal", decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "cmvae_masked": dict(model="cmvae", dims=_DIMS3, B=6, masks=True, cfg=dict(K=4, latent_dim=5, modalities_specific_dim=3, beta=2.5, loss="dreg_looser", number_of_clusters=4, prior_and_posterior_dist="laplace_with_softmax", decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    "crmvae": dict(model="crmvae", dims=_DIMS3, B=6, cfg=dict(latent_dim=5, beta=2.5, decoders_dist=_DIST3, decoder_dist_params=_PAR3, uses_likelihood_rescaling=True)),
    "crmvae_masked": dict(model="crmvae", dims=_DIMS3, B=6, masks=True, cfg=dict(latent_dim=5, beta=1.0, decoders_dist=_DIST3, decoder_dist_params=_PAR3)),
    # ---- BASELINE.json configs at their real input shapes (default MLP architectures; every hyper-parameter written out — the values
    # are the reference configs' defaults / the example scripts' settings; small batches where the config's own is big)
    "cfg1_mvtcae_quickstart": dict(model="mvtcae", dims={"mnist": (1, 28, 28), "svhn": (3, 32, 32)}, B=32,
                                   cfg=dict(latent_dim=20, alpha=0.1, beta=2.5, decoders_dist={"mnist": "normal", "svhn": "normal"},
                                            decoder_dist_params={"mnist": {}, "svhn": {}})),
    # cfg2-cfg4 on the networks SURVEY 8(d) names (mnist: default MLP, svhn: Encoder/Decoder_VAE_SVHN, PolyMNIST:
    # EncoderConvMMNIST_adapted / DecoderConvMMNIST) at the configurations' own per-device batch sizes
    "cfg2_mvae_mnistsvhn": dict(model="mvae", dims={"mnist": (1, 28, 28), "svhn": (3, 32, 32)}, B=512, fwd=dict(epoch=12, batch_ratio=0.5),
                                arch={"mnist": "mlp", "svhn": "svhn"},
                                cfg=dict(latent_dim=20, k=0, beta=1.0, warmup=10, uses_likelihood_rescaling=True,
                                         decoders_dist={"mnist": "normal", "svhn": "normal"}, decoder_dist_params={"mnist": {}, "svhn": {}})),
    "cfg3_mmvae_mnistsvhn": dict(model="mmvae", dims={"mnist": (1, 28, 28), "svhn": (3, 32, 32)}, B=256,
                                 arch={"mnist": "mlp", "svhn": "svhn"},
                                 cfg=dict(K=10, latent_dim=20, loss="iwae_looser", prior_and_posterior_dist="laplace_with_softmax",
                                          decoders_dist={"mnist": "laplace", "svhn": "laplace"},
                                          decoder_dist_params={"mnist": {"scale": 0.75}, "svhn": {"scale": 0.75}})),
    "cfg4_mopoe_polymnist": dict(model="mopoe", dims={f"m{i}": (3, 28, 28) for i in range(5)}, B=256,
                                 arch={f"m{i}": "conv_mmnist" for i in range(5)},
                                 cfg=dict(latent_dim=512, beta=2.5, decoders_dist={f"m{i}": "laplace" for i in range(5)},
                                          decoder_dist_params={f"m{i}": {"scale": 0.75} for i in range(5)})),
    "cfg5_mmvaeplus_celeba": dict(model="mmvaeplus", dims={"image": (3, 64, 64), "attributes": (40,)}, B=4,
                                  cfg=dict(K=10, latent_dim=32, modalities_specific_dim=32, beta=1.0, loss="dreg_looser",
                                           prior_and_posterior_dist="laplace_with_softmax",
                                           decoders_dist={"image": "normal", "attributes": "normal"},
                                           decoder_dist_params={"image": {}, "attributes": {}})),
    "cfg5_mmvaeplus_celeba_b128": dict(model="mmvaeplus", dims={"image": (3, 64, 64), "attributes": (40,)}, B=128,
                                       cfg=dict(K=10, latent_dim=32, modalities_specific_dim=32, beta=2.5, loss="dreg_looser",
                                                prior_and_posterior_dist="laplace_with_softmax",
                                                decoders_dist={"image": "normal", "attributes": "normal"},
                                                decoder_dist_params={"image": {}, "attributes": {}})),
    # the north star: MMVAE+ PolyMNIST, 5 modalities, K = 10, DReG, EncoderResnetMMNIST(32, 32) / DecoderResnetMMNIST(64)
    # (examples/case_studies/mmvaePlus_on_partial_data/train.py:49-78), synthetic (non-initial) weights
    "ns_mmvaeplus_resnet": dict(model="mmvaeplus", dims={f"m{i}": (3, 28, 28) for i in range(5)}, B=4,
                                arch={f"m{i}": "resnet_mmnist" for i in range(5)},
                                cfg=dict(K=10, latent_dim=32, modalities_specific_dim=32, beta=2.5, loss="dreg_looser",
                                         prior_and_posterior_dist="laplace_with_softmax", learn_modality_prior=True,
                                         learn_shared_prior=False,
                                         decoders_dist={f"m{i}": "laplace" for i in range(5)},
                                         decoder_dist_params={f"m{i}": {"scale": 0.75} for i in range(5)})),
}


def make_data(spec):
    """Synthetic inputs: U[0,1) from per-modality generators."""
    data = {}
    for i, (m, d) in enumerate(spec["dims"].items()):
        data[m] = torch.rand(spec["B"], *d, generator=torch.Generator().manual_seed(1000 + i))
        if m in spec.get("onehot", ()):   # categorical targets: one-hot over the last dimension
            data[m] = torch.nn.functional.one_hot(data[m].argmax(-1), d[-1]).float()
        if spec["cfg"].get("decoders_dist", {}).get(m) == "bernoulli":
            data[m] = (data[m] > 0.5).float()
    masks = None
    if spec.get("masks"):
        B = spec["B"]
        mods = list(spec["dims"])
        masks = {m: torch.ones(B, dtype=torch.bool) for m in mods}
        masks[mods[1]][: B // 2] = False  # mod1 missing for the first half (like the reference's test fixture)
        masks[mods[2]][B - 1] = False
        masks[mods[0]][B // 2] = False
    return data, masks
