"""Family B: mmsegmentation's own UNet-S5-D16 + FCNHead behind the same contract (SURVEY.md §0.1, §8a last rows):

  BasicConvBlock  <- mmseg/models/backbones/unet.py:16-86
  DeconvModule    <- unet.py:89-147           (ConvTranspose2d(4,2,1) upsampler: not built yet -> NotImplementedError)
  InterpConv      <- unet.py:150-221
  UNet            <- unet.py:224-438
  UpConvBlock     <- mmseg/models/utils/up_conv_block.py:9-102
  FCNHead         <- mmseg/models/decode_heads/fcn_head.py:10-88
  ConvModule      <- mmcv.cnn.ConvModule (ext): conv (bias=False when a norm follows) -> norm -> act, children `conv`/`bn`/`activate`

Same kernels as family A: tcgen05 implicit-GEMM convs + BN/ReLU, MaxPool, bilinear x2 (align_corners=False), channel concat,
classifier + fused loss.  Unsupported corners of the reference's argument space (stride-2 convs, dilation > 1, with_cp,
dcn, plugins, GroupNorm, DeconvModule) raise instead of silently doing something else.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .modules import BaseDecodeHead, _to_nchw_view, _to_nhwc, _DTYPES
from .registry import BACKBONES, HEADS, BaseModule


class ConvModule(nn.Module):
    """Parameter container with mmcv.cnn.ConvModule's child names; order conv -> norm -> act."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias="auto",
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True):
        super().__init__()
        if conv_cfg not in (None, dict(type="Conv2d")) and not (isinstance(conv_cfg, dict) and conv_cfg.get("type") in ("Conv2d", "Conv")):
            raise NotImplementedError(f"conv_cfg {conv_cfg} is not on the UNet path")
        if stride != 1 or dilation != 1 or groups != 1:
            raise NotImplementedError("stride/dilation/groups != 1 are not built (UNet-S5-D16 uses MaxPool downsampling, dilation 1)")
        if padding != (kernel_size // 2):
            raise NotImplementedError("only 'same' padding (k//2) is supported")
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias)
        if self.with_norm:
            typ = norm_cfg.get("type")
            if typ not in ("BN", "SyncBN", "BN2d"):
                raise NotImplementedError(f"norm type {typ} is not on the UNet path")
            self.bn = nn.SyncBatchNorm(out_channels) if typ == "SyncBN" else nn.BatchNorm2d(out_channels)
            for p in self.bn.parameters():
                p.requires_grad = norm_cfg.get("requires_grad", True)
        if self.with_activation:
            if act_cfg.get("type") != "ReLU":
                raise NotImplementedError(f"activation {act_cfg} is not on the UNet path")
            self.activate = nn.ReLU(inplace=inplace)

    def forward(self, x):  # NHWC
        act = ACT_RELU if self.with_activation else ACT_NONE
        if self.with_norm:
            return ops.conv_bn_act(x, self.conv, self.bn, act, self.bn.training)
        return ops.conv2d(x, self.conv.weight, self.conv.bias, act=act)

    def forward_cat(self, a, b):
        """forward(torch.cat([a, b], dim=1)): the conv reads both tensors in place when the tcgen05 kernels take them (virtual concat)."""
        if self.with_norm:
            act = ACT_RELU if self.with_activation else ACT_NONE
            return ops.conv_bn_act_cat([a, b], self.conv, self.bn, act, self.bn.training, materialise=lambda xs: ops.concat_channels(*xs))
        return self.forward(ops.concat_channels(a, b))


class BasicConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, num_convs=2, stride=1, dilation=1, with_cp=False, conv_cfg=None,
                 norm_cfg=dict(type="BN"), act_cfg=dict(type="ReLU"), dcn=None, plugins=None):
        super().__init__()
        assert dcn is None, "Not implemented yet."
        assert plugins is None, "Not implemented yet."
        if with_cp:
            raise NotImplementedError("with_cp (activation checkpointing) is not built")
        self.with_cp = with_cp
        self.convs = nn.Sequential(*[
            ConvModule(in_channels if i == 0 else out_channels, out_channels, kernel_size=3, stride=stride if i == 0 else 1,
                       dilation=1 if i == 0 else dilation, padding=1 if i == 0 else dilation, conv_cfg=conv_cfg,
                       norm_cfg=norm_cfg, act_cfg=act_cfg) for i in range(num_convs)])

    def forward(self, x):
        for m in self.convs:
            x = m(x)
        return x

    def forward_cat(self, a, b):
        x = self.convs[0].forward_cat(a, b)
        for m in list(self.convs)[1:]:
            x = m(x)
        return x


class InterpConv(nn.Module):
    def __init__(self, in_channels, out_channels, with_cp=False, norm_cfg=dict(type="BN"), act_cfg=dict(type="ReLU"), *,
                 conv_cfg=None, conv_first=False, kernel_size=1, stride=1, padding=0,
                 upsample_cfg=dict(scale_factor=2, mode="bilinear", align_corners=False)):
        super().__init__()
        if with_cp:
            raise NotImplementedError("with_cp is not built")
        if upsample_cfg.get("scale_factor", 2) != 2 or upsample_cfg.get("mode", "bilinear") != "bilinear":
            raise NotImplementedError("InterpConv: only bilinear x2 upsampling is built")
        self.align_corners = bool(upsample_cfg.get("align_corners", False))
        self.conv_first = conv_first
        conv = ConvModule(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, conv_cfg=conv_cfg,
                          norm_cfg=norm_cfg, act_cfg=act_cfg)
        upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=self.align_corners)
        self.interp_upsample = nn.Sequential(conv, upsample) if conv_first else nn.Sequential(upsample, conv)

    def forward(self, x):
        if self.conv_first:
            return ops.upsample2x(self.interp_upsample[0](x), self.align_corners)
        return self.interp_upsample[1](ops.upsample2x(x, self.align_corners))


class DeconvModule(nn.Module):
    """unet.py:89-147: ConvTranspose2d(kernel_size, stride=scale_factor, padding=(k-s)/2) -> norm -> act; state_dict keys
    `deconv_upsamping.{0,1}.*` as in the reference.  Built for the default geometry (kernel 4, scale 2)."""

    def __init__(self, in_channels, out_channels, with_cp=False, norm_cfg=dict(type="BN"), act_cfg=dict(type="ReLU"), *, kernel_size=4,
                 scale_factor=2):
        super().__init__()
        assert (kernel_size - scale_factor >= 0) and (kernel_size - scale_factor) % 2 == 0, (
            f"kernel_size should be greater than or equal to scale_factor and (kernel_size - scale_factor) should be even numbers, "
            f"while the kernel size is {kernel_size} and scale_factor is {scale_factor}.")
        if (kernel_size, scale_factor) != (4, 2):
            raise NotImplementedError("DeconvModule: only kernel_size=4, scale_factor=2 (the reference default) has a CUDA path")
        self.with_cp = with_cp
        deconv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=scale_factor,
                                    padding=(kernel_size - scale_factor) // 2)
        assert norm_cfg is not None and norm_cfg.get("type") in ("BN", "SyncBN"), "DeconvModule: BN / SyncBN only"
        norm = (nn.SyncBatchNorm if norm_cfg["type"] == "SyncBN" else nn.BatchNorm2d)(out_channels)
        for p in norm.parameters():
            p.requires_grad = norm_cfg.get("requires_grad", True)
        assert act_cfg is None or act_cfg.get("type") == "ReLU", "DeconvModule: ReLU only"
        self.act = ACT_NONE if act_cfg is None else ACT_RELU
        self.deconv_upsamping = nn.Sequential(deconv, norm, nn.ReLU(inplace=True) if act_cfg is not None else nn.Identity())

    def forward(self, x):
        deconv, norm = self.deconv_upsamping[0], self.deconv_upsamping[1]
        return ops.bn_act(ops.deconv4x2(x, deconv.weight, deconv.bias, bias_feeds_train_bn=norm.training), norm, self.act, norm.training)


_UPSAMPLE = {"InterpConv": InterpConv, "DeconvModule": DeconvModule}


class UpConvBlock(nn.Module):
    def __init__(self, conv_block, in_channels, skip_channels, out_channels, num_convs=2, stride=1, dilation=1, with_cp=False,
                 conv_cfg=None, norm_cfg=dict(type="BN"), act_cfg=dict(type="ReLU"), upsample_cfg=dict(type="InterpConv"),
                 dcn=None, plugins=None):
        super().__init__()
        assert dcn is None, "Not implemented yet."
        assert plugins is None, "Not implemented yet."
        self.conv_block = conv_block(in_channels=2 * skip_channels, out_channels=out_channels, num_convs=num_convs, stride=stride,
                                     dilation=dilation, with_cp=with_cp, conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg,
                                     dcn=None, plugins=None)
        if upsample_cfg is not None:
            cfg = dict(upsample_cfg)
            typ = cfg.pop("type")
            if typ not in _UPSAMPLE:
                raise KeyError(f"{typ} is not in the upsample registry")
            self.upsample = _UPSAMPLE[typ](in_channels=in_channels, out_channels=skip_channels, with_cp=with_cp, norm_cfg=norm_cfg,
                                           act_cfg=act_cfg, **cfg)
        else:
            self.upsample = ConvModule(in_channels, skip_channels, kernel_size=1, stride=1, padding=0, conv_cfg=conv_cfg,
                                       norm_cfg=norm_cfg, act_cfg=act_cfg)

    def forward(self, skip, x):
        x = self.upsample(x)
        return self.conv_block.forward_cat(skip, x)      # up_conv_block.py:99 torch.cat([skip, x], dim=1), never materialised


@BACKBONES.register_module()
class UNet(BaseModule):
    def __init__(self, in_channels=3, base_channels=64, num_stages=5, strides=(1, 1, 1, 1, 1), enc_num_convs=(2, 2, 2, 2, 2),
                 dec_num_convs=(2, 2, 2, 2), downsamples=(True, True, True, True), enc_dilations=(1, 1, 1, 1, 1),
                 dec_dilations=(1, 1, 1, 1), with_cp=False, conv_cfg=None, norm_cfg=dict(type="BN"), act_cfg=dict(type="ReLU"),
                 upsample_cfg=dict(type="InterpConv"), norm_eval=False, dcn=None, plugins=None, pretrained=None, init_cfg=None,
                 compute_dtype="bf16"):
        super().__init__(init_cfg)
        assert not (init_cfg and pretrained), "init_cfg and pretrained cannot be setting at the same time"
        if pretrained is not None:
            raise NotImplementedError("pretrained checkpoints are loaded through load_state_dict")
        assert dcn is None, "Not implemented yet."
        assert plugins is None, "Not implemented yet."
        for name, val, n in (("strides", strides, num_stages), ("enc_num_convs", enc_num_convs, num_stages),
                             ("dec_num_convs", dec_num_convs, num_stages - 1), ("downsamples", downsamples, num_stages - 1),
                             ("enc_dilations", enc_dilations, num_stages), ("dec_dilations", dec_dilations, num_stages - 1)):
            assert len(val) == n, f"The length of {name} should be equal to {n}, while the {name} is {val}"
        self.num_stages, self.strides, self.downsamples = num_stages, strides, downsamples
        self.norm_eval, self.base_channels = norm_eval, base_channels
        self.encoder, self.decoder = nn.ModuleList(), nn.ModuleList()
        for i in range(num_stages):
            enc = []
            if i != 0:
                if strides[i] == 1 and downsamples[i - 1]:
                    enc.append(nn.MaxPool2d(kernel_size=2))
                upsample = strides[i] != 1 or downsamples[i - 1]
                self.decoder.append(UpConvBlock(conv_block=BasicConvBlock, in_channels=base_channels * 2 ** i,
                                                skip_channels=base_channels * 2 ** (i - 1), out_channels=base_channels * 2 ** (i - 1),
                                                num_convs=dec_num_convs[i - 1], stride=1, dilation=dec_dilations[i - 1], with_cp=with_cp,
                                                conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg,
                                                upsample_cfg=upsample_cfg if upsample else None, dcn=None, plugins=None))
            enc.append(BasicConvBlock(in_channels=in_channels, out_channels=base_channels * 2 ** i, num_convs=enc_num_convs[i],
                                      stride=strides[i], dilation=enc_dilations[i], with_cp=with_cp, conv_cfg=conv_cfg,
                                      norm_cfg=norm_cfg, act_cfg=act_cfg, dcn=None, plugins=None))
            self.encoder.append(nn.Sequential(*enc))
            in_channels = base_channels * 2 ** i
        self.compute_dtype = _DTYPES[compute_dtype] if isinstance(compute_dtype, str) else compute_dtype

    def init_weights(self):
        """mmcv init_cfg default of UNet: Kaiming for Conv2d, constant 1 for norm layers (unet.py:316-323)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, a=0, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.modules.batchnorm._BatchNorm):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _check_input_divisible(self, x):
        # decoded uint8 pixels and pre-normalised device batches arrive as (N, H, W, C); the reference interface is (N, C, H, W)
        nhwc = x.dtype == torch.uint8 or isinstance(x, ops.NHWCImage)
        h, w = (x.shape[1], x.shape[2]) if nhwc else x.shape[-2:]
        rate = 1
        for i in range(1, self.num_stages):
            if self.strides[i] == 2 or self.downsamples[i - 1]:
                rate *= 2
        assert h % rate == 0 and w % rate == 0, \
            f"The input image size {(h, w)} should be divisible by the whole downsample rate {rate}, when num_stages is " \
            f"{self.num_stages}, strides is {self.strides}, and downsamples is {self.downsamples}."

    def forward(self, x):
        self._check_input_divisible(x)
        h = ops.image_to_nhwc(x, self.compute_dtype, getattr(self, "img_norm_cfg", None))
        enc_outs = []
        last = len(self.encoder) - 1
        for i, enc in enumerate(self.encoder):
            for m in enc:
                h = ops.maxpool2(h) if isinstance(m, nn.MaxPool2d) else m(h)
            if i < last:   # feeds both the next stage and a decoder skip: explicit fan-out (our add kernel sums the grads)
                h, skip = ops.fanout(h, 2)
                enc_outs.append(skip)
        # every map is returned to the heads AND (except the final one) consumed by the next decoder block
        h, ret = ops.fanout(h, 2)
        dec_outs = [ret]
        for i in reversed(range(len(self.decoder))):
            h = self.decoder[i](enc_outs[i], h)
            if i > 0:
                h, ret = ops.fanout(h, 2)
                dec_outs.append(ret)
            else:
                dec_outs.append(h)
        return [_to_nchw_view(t) for t in dec_outs]

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        return self


@HEADS.register_module()
class FCNHead(BaseDecodeHead):
    def __init__(self, num_convs=2, kernel_size=3, concat_input=True, dilation=1, **kwargs):
        assert num_convs >= 0 and dilation > 0 and isinstance(dilation, int)
        self.num_convs, self.concat_input, self.kernel_size = num_convs, concat_input, kernel_size
        super().__init__(**kwargs)
        if num_convs == 0:
            assert self.in_channels == self.channels
        if dilation != 1:
            raise NotImplementedError("FCNHead dilation > 1 is not built")
        convs = [ConvModule(self.in_channels if i == 0 else self.channels, self.channels, kernel_size=kernel_size,
                            padding=kernel_size // 2, conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg, act_cfg=self.act_cfg)
                 for i in range(num_convs)]
        self.convs = nn.Sequential(*convs) if convs else nn.Identity()
        if self.concat_input:
            self.conv_cat = ConvModule(self.in_channels + self.channels, self.channels, kernel_size=kernel_size,
                                       padding=kernel_size // 2, conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg, act_cfg=self.act_cfg)

    def _forward_feature(self, inputs):
        x = ops._chk(_to_nhwc(inputs[self.in_index]))
        feats = x
        if self.num_convs > 0:
            xa, xb = ops.fanout(x, 2) if self.concat_input else (x, x)
            feats = xa
            for m in self.convs:
                feats = m(feats)
            if self.concat_input:
                feats = self.conv_cat.forward_cat(xb, feats)
        elif self.concat_input:
            xa, xb = ops.fanout(x, 2)
            feats = self.conv_cat.forward_cat(xb, xa)
        return feats

    def forward(self, inputs):
        return self.cls_seg(self._forward_feature(inputs))
