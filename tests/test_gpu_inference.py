"""Inference API and likelihood estimators on the GPU against goldens of the real reference (see tests/infer_checks.py)."""
import pytest
import torch

from oracle.make_golden_infer import INFER_CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", INFER_CASES)
def test_inference_api_matches_reference(name):
    from tests.infer_checks import check_infer_case
    check_infer_case(name)


def test_eval_step_matches_train_loss_conventions():
    """eval_step: no_grad forward over the eval set, loss_sum / len(dataset), metrics averaged over batches (base_trainer.py:618-680)."""
    import multivae_b200 as mb
    from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig
    torch.manual_seed(0)
    dims = {"a": (3, 8, 8), "b": (10,)}
    model = mb.MVTCAE(mb.MVTCAEConfig(n_modalities=2, latent_dim=6, input_dims=dict(dims)))
    data = {m: torch.rand(24, *d) for m, d in dims.items()}
    tr = BaseTrainer(model, mb.MultimodalBaseDataset(data=data), eval_dataset=mb.MultimodalBaseDataset(data=data),
                     training_config=BaseTrainerConfig(per_device_train_batch_size=8, per_device_eval_batch_size=12))
    before = {k: v.clone() for k, v in model.state_dict().items()}
    loss, metrics = tr.eval_step(epoch=1)
    assert isinstance(loss, float) and loss == loss and set(metrics) >= {"joint_divergence", "a", "b"}
    assert all(torch.equal(before[k], v) for k, v in model.state_dict().items())
    assert all(p.grad is None or float(p.grad.abs().sum()) == 0.0 for p in model.parameters())


def test_likelihoods_evaluator_matches_the_golden_joint_nll():
    """LikelihoodsEvaluator (metrics/likelihoods/likelihoods.py:43-61) over the golden case's data as ONE batch with the golden's
    recorded noise = the reference's `compute_joint_nll` value / n_data; in two batches with fresh noise it stays within Monte-Carlo
    distance of it."""
    import os

    import multivae_b200 as mb
    from oracle.cases import CASES, make_data
    from oracle.make_golden_infer import NLL_BK, NLL_K
    from tests.infer_checks import GOLD, _feed, _model
    name = "mvtcae"
    rec = torch.load(os.path.join(GOLD, f"infer_{name}.pt"), weights_only=False)
    spec = CASES[name]
    model = _model(name, rec, "cuda")
    data, _ = make_data(spec)
    ds = mb.MultimodalBaseDataset(data=data)
    cfg = mb.LikelihoodsEvaluatorConfig(batch_size=spec["B"], num_samples=NLL_K, batch_size_k=NLL_BK)
    ev = mb.LikelihoodsEvaluator(model, ds, eval_config=cfg)
    q = _feed(model, rec["calls"]["joint_nll"]["noise"], "cuda")
    out = ev.eval()
    assert not q
    ref = float(rec["calls"]["joint_nll"]["out"]) / spec["B"]
    assert abs(float(out.joint_likelihood) - ref) <= 2e-4 * abs(ref)
    model.noise_source = None
    torch.manual_seed(0)
    ev2 = mb.LikelihoodsEvaluator(model, ds, eval_config=mb.LikelihoodsEvaluatorConfig(batch_size=spec["B"] // 2, num_samples=400, batch_size_k=100))
    assert abs(float(ev2.eval().joint_likelihood) - ref) <= 0.05 * abs(ref)
