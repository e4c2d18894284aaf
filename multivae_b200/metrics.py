"""Likelihood evaluation module (reference: metrics/likelihoods/likelihoods.py:13-84, metrics/base/evaluator_class.py:15-60,
metrics/likelihoods/likelihoods_config.py): walks a test set in batches and averages the models' importance-sampled joint
negative log-likelihoods.  The estimators themselves (`compute_joint_nll`, `compute_joint_nll_paper`,
`_compute_joint_nll_from_subset_encoding`) are the batched native ones of the models; this class is the reference's loop around
them.  The reference's wandb hook and the other evaluators (coherence, FID, clustering: they need trained classifiers /
Inception weights) are out of scope."""
import logging
from dataclasses import dataclass
from pathlib import Path

import torch

from .containers import ModelOutput, MultimodalBaseDataset, set_inputs_to_device


@dataclass
class EvaluatorConfig:
    batch_size: int = 512
    wandb_path: str = None


@dataclass
class LikelihoodsEvaluatorConfig(EvaluatorConfig):
    num_samples: int = 1000
    batch_size_k: int = 100
    unified_implementation: bool = True


class Evaluator:
    """Base class (evaluator_class.py:15-60): model to the device in eval mode, the test set cut into batches in order."""

    def __init__(self, model, test_dataset, output=None, eval_config=None, sampler=None):
        eval_config = EvaluatorConfig() if eval_config is None else eval_config
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.model = model.to(self.device).eval()
        self.n_data = len(test_dataset)
        self.batch_size = eval_config.batch_size
        self.test_dataset = test_dataset
        if output is not None:
            Path(output).mkdir(parents=True, exist_ok=True)
        self.output = output
        self.logger = logging.getLogger(f"multivae_b200.metrics.{id(self)}")
        self.logger.setLevel(logging.INFO)
        if output is not None:
            self.logger.addHandler(logging.FileHandler(str(output) + "/metrics.log"))
        self.metrics = {}
        self.sampler = sampler
        if sampler is not None and not sampler.is_fitted:
            raise AttributeError("The provided sampler is not fitted.Please fit the sampler before using it in the evaluator module.")

    @property
    def test_loader(self):
        """Batches of the test set in dataset order (the reference's `DataLoader(test_dataset, batch_size)`)."""
        ds = self.test_dataset
        for i0 in range(0, self.n_data, self.batch_size):
            idx = slice(i0, min(i0 + self.batch_size, self.n_data))
            yield set_inputs_to_device(ds[idx], self.device)


class LikelihoodsEvaluator(Evaluator):
    """likelihoods.py:13-84."""

    def __init__(self, model, test_dataset, output=None, eval_config=None):
        eval_config = LikelihoodsEvaluatorConfig() if eval_config is None else eval_config
        super().__init__(model, test_dataset, output, eval_config)
        self.num_samples = eval_config.num_samples
        self.batch_size_k = eval_config.batch_size_k
        self.unified = eval_config.unified_implementation

    def eval(self):
        self.joint_nll()
        return ModelOutput(**self.metrics)

    def _batch(self, b):
        if hasattr(b, "masks"):   # the estimators refuse incomplete data, like the reference's
            from .containers import IncompleteDataset
            return IncompleteDataset(data=dict(b.data), masks=dict(b.masks))
        return MultimodalBaseDataset(data=dict(b.data))

    def joint_nll(self):
        ll = 0
        for batch in self.test_loader:
            if self.unified or not hasattr(self.model, "compute_joint_nll_paper"):
                ll = ll + self.model.compute_joint_nll(self._batch(batch), self.num_samples, self.batch_size_k)
            else:
                self.logger.info("Using the paper version of the joint nll.")
                ll = ll + self.model.compute_joint_nll_paper(self._batch(batch), self.num_samples, self.batch_size_k)
        joint_nll = ll / self.n_data
        self.logger.info(f"Mean Joint likelihood : {str(joint_nll)}")
        self.metrics["joint_likelihood"] = joint_nll
        return joint_nll

    def joint_nll_from_subset(self, subset):
        """MoPoE only: a subset posterior as the importance distribution (likelihoods.py:63-84)."""
        if not hasattr(self.model, "_compute_joint_nll_from_subset_encoding"):
            return None
        ll = 0
        for batch in self.test_loader:
            ll = ll + self.model._compute_joint_nll_from_subset_encoding(subset, self._batch(batch), self.num_samples, self.batch_size_k)
        joint_nll = ll / self.n_data
        self.logger.info("Joint likelihood from subset %s", str(joint_nll))
        self.metrics[f"Joint likelihood from subset {subset}"] = joint_nll
        return joint_nll
