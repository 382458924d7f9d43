"""stc_unet_b200 — B200-native (sm_100a) implementation of the STC-UNet training / inference hot path.

Importing the package registers `UnetBackbone`, `UnetHead`, `CrossEntropyLoss`, `DiceLoss` and
`EncoderDecoder` (see registry.py for how they plug into mmsegmentation's registry).  All compute
runs in libstc_b200.so (hand-written CUDA, C ABI in include/stc_b200.h); there is no CPU fallback.
"""
from . import _lib, ops, registry  # noqa: F401
from .modules import (BaseDecodeHead, CoordAtt, CrossEntropyLoss, DiceLoss, DoubleConv, Down, InConv,  # noqa: F401
                      KernelSelectAttention, TransformerBlock, TransformerLayer, UnetBackbone, UnetHead, Up)
from .modules_b import BasicConvBlock, ConvModule, FCNHead, InterpConv, UNet, UpConvBlock  # noqa: F401
from .modules_pp import EncoderDecoderFull, UnetPlusPlus  # noqa: F401
from .registry import BACKBONES, HEADS, LOSSES, MODELS, SEGMENTORS, build_backbone, build_head, build_loss, build_segmentor  # noqa: F401
from .segmentor import EncoderDecoder, slide_windows  # noqa: F401
from .optim import OPTIMIZERS, AdamB200, PolyLrUpdater, build_optimizer, load_checkpoint, poly_lr, save_checkpoint  # noqa: F401

__version__ = "0.1.0"
