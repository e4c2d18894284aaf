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
