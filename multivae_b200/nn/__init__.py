"""Encoder / decoder architectures with the reference's parameter names and shapes
(/root/reference/src/multivae/models/nn/*.py), so that reference checkpoints load unchanged."""
from .base_architectures import BaseDecoder, BaseEncoder  # noqa: F401
from .default_architectures import (  # noqa: F401
    BaseDictDecoders,
    BaseDictDecodersMultiLatents,
    BaseDictEncoders,
    BaseDictEncoders_MultiLatents,
    Decoder_AE_MLP,
    Encoder_VAE_MLP,
    Encoder_VAE_MLP_Style,
)
from .mmnist import (  # noqa: F401
    DecoderConvMMNIST,
    DecoderResnetMMNIST,
    EncoderConvMMNIST_adapted,
    EncoderResnetMMNIST,
    ResnetBlock,
)
from .cub import CUB_Resnet_Decoder, CUB_Resnet_Encoder, CUB_ResnetBlock  # noqa: F401
from .svhn import Decoder_VAE_SVHN, Encoder_VAE_SVHN  # noqa: F401
