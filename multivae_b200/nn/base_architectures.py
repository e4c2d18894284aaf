"""Base classes every encoder / decoder handed to a model must derive from (the reference checks
`issubclass(type(encoder), BaseEncoder)`, models/base/base_ae_model.py:345-372)."""
import torch.nn as nn


class BaseEncoder(nn.Module):
    """forward(x) -> ModelOutput(embedding, log_covariance[, style_embedding, style_log_covariance])"""

    def forward(self, x):
        raise NotImplementedError()


class BaseDecoder(nn.Module):
    """forward(z) -> ModelOutput(reconstruction)"""

    def forward(self, z):
        raise NotImplementedError()
