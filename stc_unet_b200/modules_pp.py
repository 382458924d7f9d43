"""Config 5 (my_config/UNet++.py): `EncoderDecoderFull` + `UnetPlusPlus` head.

  UnetPlusPlus        <- mmseg/models/decode_heads/unetpp_head.py:11-22: smp.UnetPlusPlus(encoder_name="vgg16", classes=64) + cls_seg
  EncoderDecoderFull  <- mmseg/models/segmentors/encoder_decoder.py:334-583 (backbone-less EncoderDecoder: the head eats the image)

`segmentation_models_pytorch==0.2.0` is NOT vendored in the reference and not installed here, so the architecture below is a
restatement of its published design (PARITY UNPINNED): torchvision VGG16 `features` (13 conv3x3+bias+ReLU, no BN,
split at every MaxPool into 6 stages: 64, 128, 256, 512, 512 channels and a final pool-only stage), the UNet++ nested decoder
(blocks x_d_l = nearest x2 -> cat([x, dense skips.., encoder feature]) -> 2 x [conv3x3(no bias) -> BN -> ReLU], decoder
channels (256,128,64,32,16), the full-resolution 64-channel encoder feature is NOT used, no center block in forward), and a
3x3 segmentation head 16 -> 64.  state_dict keys follow smp: `model.encoder.features.<i>.*`, `model.decoder.blocks.x_<d>_<l>.
conv{1,2}.{0,1}.*`, `model.segmentation_head.0.*`, plus `conv_seg.*`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .modules import BaseDecodeHead, _DTYPES
from .registry import HEADS, SEGMENTORS
from .segmentor import EncoderDecoder

_VGG16 = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M")


class _VGGEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        layers, c = [], 3
        for v in _VGG16:
            if v == "M":
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                layers += [nn.Conv2d(c, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
                c = v
        self.features = nn.Sequential(*layers)

    def forward(self, x):  # NHWC -> 6 features (stage outputs)
        feats = []
        for m in self.features:
            if isinstance(m, nn.MaxPool2d):
                x, keep = ops.fanout(x, 2)     # stage output: read by the decoder and by the next stage's pool
                feats.append(keep)
                x = ops.maxpool2(x)
            elif isinstance(m, nn.Conv2d):
                x = ops.conv2d(x, m.weight, m.bias, act=ACT_RELU)
        feats.append(x)
        return feats


class _Conv2dReLU(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x):
        return ops.conv_bn_act(x, self[0], self[1], ACT_RELU, self[1].training)


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _Conv2dReLU(cin + cskip, cout)
        self.attention1 = nn.Identity()
        self.conv2 = _Conv2dReLU(cout, cout)
        self.attention2 = nn.Identity()

    def forward(self, x, skips=()):
        skips = list(skips)
        chs = [x.shape[-1]] + [t.shape[-1] for t in skips]
        if skips and ops.cat_ok(chs, self.conv1[0].weight.shape[0], x.dtype):
            # virtual concat: only the nearest-x2 map is written; the dense skips are read in place by conv1's K loop
            xu = ops.cat_channels_n([x], up0=True)
            y = ops.conv_bn_act_cat([xu, *skips], self.conv1[0], self.conv1[1], ACT_RELU, self.conv1[1].training)
            return self.conv2(y)
        x = ops.cat_channels_n([x, *skips], up0=True)
        return self.conv2(self.conv1(x))


class _UnetPlusPlusDecoder(nn.Module):
    def __init__(self, encoder_channels=(64, 128, 256, 512, 512, 512), decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        head = enc[0]
        self.in_channels = [head] + list(decoder_channels[:-1])
        self.skip_channels = list(enc[1:]) + [0]
        self.out_channels = list(decoder_channels)
        blocks = {}
        for l in range(len(self.in_channels) - 1):
            for d in range(l + 1):
                if d == 0:
                    cin, cskip, cout = self.in_channels[l], self.skip_channels[l] * (l + 1), self.out_channels[l]
                else:
                    cout, cskip, cin = self.skip_channels[l], self.skip_channels[l] * (l + 1 - d), self.skip_channels[l - 1]
                blocks[f"x_{d}_{l}"] = _DecoderBlock(cin, cskip, cout)
        blocks[f"x_0_{len(self.in_channels) - 1}"] = _DecoderBlock(self.in_channels[-1], 0, self.out_channels[-1])
        self.blocks = nn.ModuleDict(blocks)
        self.depth = len(self.in_channels) - 1

    def forward(self, features):
        features = features[1:][::-1]      # drop the full-resolution stage, deepest first
        uses = {}                          # every dense node / feature is read by several blocks: explicit fan-out

        def take(key, tensor=None):
            if key not in uses:
                n = self._consumers[key]
                uses[key] = list(ops.fanout(tensor, n)) if n > 1 else [tensor]
            return uses[key].pop()

        if not hasattr(self, "_consumers"):
            self._consumers = self._count_consumers()
        feats = {f"f{i}": f for i, f in enumerate(features)}
        dense = {}
        for l in range(len(self.in_channels) - 1):
            for d in range(self.depth - l):
                if l == 0:
                    out = self.blocks[f"x_{d}_{d}"](take(f"f{d}", feats[f"f{d}"]), [take(f"f{d + 1}", feats[f"f{d + 1}"])])
                    dense[f"x_{d}_{d}"] = out
                else:
                    li = d + l
                    cat = [take(f"x_{i}_{li}", dense[f"x_{i}_{li}"]) for i in range(d + 1, li + 1)] + [take(f"f{li + 1}", feats[f"f{li + 1}"])]
                    dense[f"x_{d}_{li}"] = self.blocks[f"x_{d}_{li}"](take(f"x_{d}_{li - 1}", dense[f"x_{d}_{li - 1}"]), cat)
        return self.blocks[f"x_0_{self.depth}"](take(f"x_0_{self.depth - 1}", dense[f"x_0_{self.depth - 1}"]))

    def _count_consumers(self):
        cnt = {}

        def use(k):
            cnt[k] = cnt.get(k, 0) + 1
        for l in range(len(self.in_channels) - 1):
            for d in range(self.depth - l):
                if l == 0:
                    use(f"f{d}"); use(f"f{d + 1}")
                else:
                    li = d + l
                    for i in range(d + 1, li + 1):
                        use(f"x_{i}_{li}")
                    use(f"f{li + 1}"); use(f"x_{d}_{li - 1}")
        use(f"x_0_{self.depth - 1}")
        return cnt


class _SmpUnetPlusPlus(nn.Module):
    def __init__(self, classes=64):
        super().__init__()
        self.encoder = _VGGEncoder()
        self.decoder = _UnetPlusPlusDecoder()
        self.segmentation_head = nn.Sequential(nn.Conv2d(16, classes, kernel_size=3, padding=1), nn.Identity(), nn.Identity())

    def forward(self, x):  # NHWC
        y = self.decoder(self.encoder(x))
        h = self.segmentation_head[0]
        return ops.conv2d(y, h.weight, h.bias, act=ACT_NONE)


@HEADS.register_module()
class UnetPlusPlus(BaseDecodeHead):
    def __init__(self, num_classes, deep_supervision=False, compute_dtype="bf16", **kwargs):
        super().__init__(num_classes=num_classes, **kwargs)
        self.num_classes = num_classes
        self.model = _SmpUnetPlusPlus(classes=64)
        self.compute_dtype = _DTYPES[compute_dtype] if isinstance(compute_dtype, str) else compute_dtype

    def forward(self, x):  # x: the IMAGE (N,3,H,W) — EncoderDecoderFull has no backbone
        h = ops.image_to_nhwc(x, self.compute_dtype, getattr(self, "img_norm_cfg", None))
        return self.cls_seg(self.model(h))


@SEGMENTORS.register_module()
class EncoderDecoderFull(EncoderDecoder):
    """encoder_decoder.py:334-583: the decode head is the whole network."""

    def __init__(self, decode_head, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None, **_):
        from .registry import BaseModule, build_head
        BaseModule.__init__(self, init_cfg)
        self.backbone = None
        self.decode_head = build_head(decode_head) if isinstance(decode_head, dict) else decode_head
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.out_channels = self.decode_head.out_channels
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg or dict(mode="whole")

    def extract_feat(self, img):
        return img
