"""multivae_b200 — B200-native training step for multimodal VAEs (MMVAE+, MMVAE, MoPoE, MVAE, MVTCAE, CMVAE, CRMVAE)
behind MultiVae's `Model(config, encoders, decoders).forward(inputs) -> ModelOutput` API."""
from .configs import (  # noqa: F401
    BaseAEConfig,
    BaseMultiVAEConfig,
    CMVAEConfig,
    CRMVAEConfig,
    MMVAEConfig,
    MMVAEPlusConfig,
    MoPoEConfig,
    MVAEConfig,
    MVTCAEConfig,
)
from .containers import (  # noqa: F401
    DatasetOutput,
    IncompleteDataset,
    ModelOutput,
    MultimodalBaseDataset,
    set_inputs_to_device,
)
from .base import BaseMultiVAE  # noqa: F401
from .cmvae import CMVAE  # noqa: F401
from .crmvae import CRMVAE  # noqa: F401
from .mmvae import MMVAE  # noqa: F401
from .mmvae_plus import MMVAEPlus  # noqa: F401
from .mopoe import MoPoE  # noqa: F401
from .mvae import MVAE  # noqa: F401
from .metrics import EvaluatorConfig, LikelihoodsEvaluator, LikelihoodsEvaluatorConfig  # noqa: F401
from .mvtcae import MVTCAE  # noqa: F401

__version__ = "0.1.0"
