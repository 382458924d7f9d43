"""B200-native re-implementation of the reference's hot-path modules behind the same module
contract (constructor kwargs, forward signatures, state_dict keys/shapes):

  UnetBackbone            <- mmseg/models/backbones/unet_backbone.py:15-52
  KernelSelectAttention   <- unet_backbone.py:55-99
  TransformerBlock/Layer  <- unet_backbone.py:195-246
  UnetHead / Up / CoordAtt<- mmseg/models/decode_heads/unet_head.py:11-146
  BaseDecodeHead          <- mmseg/models/decode_heads/decode_head.py:15-296
  CrossEntropyLoss/DiceLoss <- mmseg/models/losses/{cross_entropy_loss,dice_loss}.py

The torch.nn layers instantiated here (Conv2d, SyncBatchNorm, Linear, MultiheadAttention) are
parameter CONTAINERS only — they give the reference's state_dict layout and default init; their
`forward` is never called.  All arithmetic goes through stc_unet_b200.ops (our CUDA kernels).
Internally activations are NHWC; tensors crossing the module boundary are NCHW-shaped views with
channels_last strides (no copy).
"""
from __future__ import annotations

import warnings
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_HSWISH, ACT_NONE, ACT_RELU, ACT_SIGMOID, ENGINE_SIMT
from .registry import BACKBONES, HEADS, LOSSES, BaseModule, build_loss

_DTYPES = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp32": torch.float32, "float32": torch.float32}


def _to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """(N,C,H,W) view -> (N,H,W,C); free when x has channels_last strides (ours always do)."""
    return x.permute(0, 2, 3, 1)


def _to_nchw_view(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------
# encoder
# ------------------------------------------------------------------------------------------
class DoubleConv(nn.Module):
    """[Conv3x3 -> SyncBN -> ReLU] x 2 (unet_backbone.py:116-130, unet_head.py:63-77)."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_ch, out_ch, 3, padding=1), nn.SyncBatchNorm(out_ch), nn.ReLU(inplace=True),
            nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.SyncBatchNorm(out_ch), nn.ReLU(inplace=True))

    def forward(self, x):  # x NHWC
        x = ops.conv_bn_act(x, self.conv[0], self.conv[1], ACT_RELU, self.training)
        return ops.conv_bn_act(x, self.conv[3], self.conv[4], ACT_RELU, self.training)

    def forward_cat(self, xs):
        """DoubleConv(torch.cat(xs, dim=1)) with the first conv reading its sources directly (virtual concat)."""
        x = ops.conv_bn_act_cat(xs, self.conv[0], self.conv[1], ACT_RELU, self.training)
        return ops.conv_bn_act(x, self.conv[3], self.conv[4], ACT_RELU, self.training)


class InConv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = DoubleConv(in_ch, out_ch)

    def forward(self, x):
        return self.conv(x)


class Down(nn.Module):
    """MaxPool2d(2) + DoubleConv (unet_backbone.py:102-113; `coord_att` is accepted and ignored, as upstream)."""

    def __init__(self, in_ch, out_ch, coord_att=False):
        super().__init__()
        self.coord_att = coord_att
        self.down_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_ch, out_ch))

    def forward(self, x):
        return self.down_conv[1](ops.maxpool2(x))


class KernelSelectAttention(nn.Module):
    def __init__(self, channel=512, kernels=(3, 5, 7), reduction=16, group=1, L=32):
        super().__init__()
        if group != 1:
            raise NotImplementedError("grouped KernelSelectAttention convs are not on the STC-UNet path")
        self.d = max(L, channel // reduction)
        self.convs = nn.ModuleList([
            nn.Sequential(nn.Conv2d(channel, channel, kernel_size=k, padding=k // 2, groups=group),
                          nn.SyncBatchNorm(channel), nn.ReLU()) for k in kernels])
        self.fc = nn.Linear(channel, self.d)
        self.fcs = nn.ModuleList([nn.Linear(self.d, channel) for _ in kernels])
        self.softmax = nn.Softmax(dim=0)

    def forward_residual(self, x):
        """Returns x + V (the caller's `x1 = x1 + res_x1`, unet_backbone.py:46-48, fused into the combine kernel)."""
        xs = ops.fanout(x, 4)
        feats = [ops.conv_bn_act(xs[i], seq[0], seq[1], ACT_RELU, self.training) for i, seq in enumerate(self.convs)]
        return ops.ksa_fuse(xs[3], feats[0], feats[1], feats[2], self.fc, self.fcs)


class TransformerLayer(nn.Module):
    def __init__(self, c, num_heads):
        super().__init__()
        self.q = nn.Linear(c, c, bias=False)
        self.k = nn.Linear(c, c, bias=False)
        self.v = nn.Linear(c, c, bias=False)
        self.ma = nn.MultiheadAttention(embed_dim=c, num_heads=num_heads)
        self.fc1 = nn.Linear(c, c, bias=False)
        self.fc2 = nn.Linear(c, c, bias=False)
        self.num_heads = num_heads

    def forward(self, t):  # t (N, L, E)
        if t.dtype == torch.bfloat16 and ops.config.fold_linear_pairs and ops.config.engine != ENGINE_SIMT:
            # same function, fewer GEMMs: (q|k|v then in_proj) and (fc1 then fc2) have nothing in between, so each pair is
            # applied as ONE folded E x E weight (ops._FusedLinearPairs); the fp32 parity path below keeps the reference's order
            t0, t3 = ops.fanout(t, 2)
            qkv = ops.fused_linear_pairs(t0, (self.q.weight, self.k.weight, self.v.weight), self.ma.in_proj_weight, self.ma.in_proj_bias)
            o = ops.attention_packed(qkv, self.num_heads)
            t = ops.linear_tokens(o, self.ma.out_proj.weight, self.ma.out_proj.bias, residual=t3)
            ta, tb = ops.fanout(t, 2)
            return ops.fused_linear_pairs(ta, (self.fc1.weight,), self.fc2.weight, None, residual=tb)
        t0, t1, t2, t3 = ops.fanout(t, 4)
        q = ops.linear_tokens(t0, self.q.weight)
        k = ops.linear_tokens(t1, self.k.weight)
        v = ops.linear_tokens(t2, self.v.weight)
        qp, kp, vp = ops.in_proj(q, k, v, self.ma)
        o = ops.attention(qp, kp, vp, self.num_heads)
        t = ops.linear_tokens(o, self.ma.out_proj.weight, self.ma.out_proj.bias, residual=t3)
        ta, tb = ops.fanout(t, 2)
        h = ops.linear_tokens(ta, self.fc1.weight)
        return ops.linear_tokens(h, self.fc2.weight, None, residual=tb)


class TransformerBlock(nn.Module):
    def __init__(self, c1, c2, num_heads, num_layers):
        super().__init__()
        if c1 != c2:
            raise NotImplementedError("TransformerBlock with c1 != c2 (the Conv+BN+SiLU stem) is not on the STC-UNet path")
        self.conv = None
        self.linear = nn.Linear(c2, c2)
        self.tr = nn.Sequential(*(TransformerLayer(c2, num_heads) for _ in range(num_layers)))
        self.c2 = c2

    def forward_residual(self, x):
        """x NHWC (N,H,W,C) -> x + block(x).  NHWC flattened IS the (N, L, E) token layout, so the reference's
        flatten/permute round trip (unet_backbone.py:245-246) costs nothing here."""
        N, H, W, C = x.shape
        xa, xb = ops.fanout(x, 2)
        t = xa.reshape(N, H * W, C)
        ta, tb = ops.fanout(t, 2)
        t = ops.linear_tokens(ta, self.linear.weight, self.linear.bias, residual=tb)
        for layer in self.tr:
            t = layer(t)
        return ops.add_autograd(t.reshape(N, H, W, C), xb)


@BACKBONES.register_module()
class UnetBackbone(BaseModule):
    def __init__(self, in_channels=3, channel_list=[64, 128, 256, 512], context_layer=None, coord_att=False,
                 transformer_block=False, compute_dtype="bf16", **kwargs):
        super().__init__(**kwargs)
        self.inc = InConv(in_channels, channel_list[0])
        self.down1 = Down(channel_list[0], channel_list[1], coord_att=coord_att)
        self.down2 = Down(channel_list[1], channel_list[2], coord_att=coord_att)
        self.down3 = Down(channel_list[2], channel_list[3], coord_att=coord_att)
        self.down4 = Down(channel_list[3], channel_list[3], coord_att=coord_att)
        self.context_layer = context_layer
        self.coord_att = coord_att
        self.transformer_block = transformer_block
        if self.context_layer == "kernelselect":
            self.context_layer1_1 = KernelSelectAttention(channel=channel_list[0])
            self.context_layer2_1 = KernelSelectAttention(channel=channel_list[1])
            self.context_layer3_1 = KernelSelectAttention(channel=channel_list[2])
        elif self.context_layer:
            raise ValueError(f"unsupported context_layer {context_layer!r}")
        if self.transformer_block:
            self.aspp4 = TransformerBlock(c1=512, c2=512, num_heads=2, num_layers=4)
            self.aspp5 = TransformerBlock(c1=512, c2=512, num_heads=2, num_layers=4)
        self.compute_dtype = _DTYPES[compute_dtype] if isinstance(compute_dtype, str) else compute_dtype

    def forward(self, x) -> List[torch.Tensor]:
        h = ops.image_to_nhwc(x, self.compute_dtype, getattr(self, "img_norm_cfg", None))  # raises on non-CUDA input: no CPU path
        x1 = self.inc(h)
        ctx, tr = bool(self.context_layer), self.transformer_block
        x1a, x1b = ops.fanout(x1, 2) if ctx else (x1, x1)
        x2 = self.down1(x1a)
        x2a, x2b = ops.fanout(x2, 2) if ctx else (x2, x2)
        x3 = self.down2(x2a)
        x3a, x3b = ops.fanout(x3, 2) if ctx else (x3, x3)
        x4 = self.down3(x3a)
        x4a, x4b = ops.fanout(x4, 2) if tr else (x4, x4)
        x5 = self.down4(x4a)
        if ctx:  # note: down_k consumed the PRE-residual maps (unet_backbone.py:37-48)
            x1 = self.context_layer1_1.forward_residual(x1b)
            x2 = self.context_layer2_1.forward_residual(x2b)
            x3 = self.context_layer3_1.forward_residual(x3b)
        if tr:
            x4 = self.aspp4.forward_residual(x4b)
            x5 = self.aspp5.forward_residual(x5)
        return [_to_nchw_view(t) for t in (x1, x2, x3, x4, x5)]


# ------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------
class h_sigmoid(nn.Module):
    def __init__(self, inplace=True):
        super().__init__()
        self.relu = nn.ReLU6(inplace=inplace)


class h_swish(nn.Module):
    def __init__(self, inplace=True):
        super().__init__()
        self.sigmoid = h_sigmoid(inplace=inplace)


class CoordAtt(nn.Module):
    """unet_head.py:116-146.  forward_add returns x + a_h * a_w (the `self.ca(x) + x` of Up.forward:57)."""

    def __init__(self, inp, oup, reduction=4):
        super().__init__()
        mip = max(8, inp // reduction)
        self.pool_h = nn.AdaptiveAvgPool2d((None, 1))
        self.pool_w = nn.AdaptiveAvgPool2d((1, None))
        self.conv1 = nn.Conv2d(inp, mip, kernel_size=1, stride=1, padding=0)
        self.bn1 = nn.SyncBatchNorm(mip)
        self.act = h_swish()
        self.conv_h = nn.Conv2d(mip, oup, kernel_size=1, stride=1, padding=0)
        self.conv_w = nn.Conv2d(mip, oup, kernel_size=1, stride=1, padding=0)

    def forward_add(self, x):  # x NHWC
        N, H, W, C = x.shape
        y, xb = ops.coordatt_pool(x)                              # (N, H+W, C) descriptors + pass-through of x
        return ops.coordatt_apply(xb, self.attention(y, N, H, W))

    def attention(self, y, N, H, W):
        """(N, H+W, C) pooled descriptors -> (N, H+W, C) attention factors [a_h ; a_w]  (unet_head.py:137-145)."""
        C = y.shape[-1]
        # conv1 + bn1 + h_swish over the N*(H+W) descriptors, as a (N, H+W, 1, C) image
        y = ops.conv_bn_act(y.view(N, H + W, 1, C), self.conv1, self.bn1, ACT_HSWISH, self.training)
        mip = y.shape[-1]
        ya, yb = ops.fanout(y, 2)
        # conv_h acts on the first H rows, conv_w on the last W rows; each gets its own 1x1 conv + sigmoid
        ah = ops.conv2d(ops.slice_rows(ya, 0, H), self.conv_h.weight, self.conv_h.bias, act=ACT_SIGMOID)
        aw = ops.conv2d(ops.slice_rows(yb, H, W), self.conv_w.weight, self.conv_w.bias, act=ACT_SIGMOID)
        a = ops.cat_rows(ah, aw)                                  # (N, H+W, 1, C)
        return a.view(N, H + W, C)


class Up(nn.Module):
    def __init__(self, in_ch, out_ch, bilinear=True, se=False):
        super().__init__()
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        else:
            raise NotImplementedError("Up(bilinear=False) (ConvTranspose2d) is never built by UnetHead")
        self.se = se
        if self.se:
            self.ca = CoordAtt(in_ch, in_ch)
        self.conv = DoubleConv(in_ch, out_ch)

    def forward(self, x1, x2):  # both NHWC; x2 = skip
        if self.se:   # cat, CoordAtt pooling and `ca(x) + x` in one fused write of cat + a_h * a_w
            x = ops.upcat_coordatt(x2, x1, True, self.ca.attention)
            return self.conv(x)
        same = x2.shape[1] == 2 * x1.shape[1] and x2.shape[2] == 2 * x1.shape[2]       # no F.pad needed (unet_head.py:52-54)
        if same and ops.cat_ok([x2.shape[-1], x1.shape[-1]], self.conv.conv[0].weight.shape[0], x1.dtype):
            # virtual concat: the up-sampled map is written once, the skip is read in place by the conv's K loop; cat never exists
            return self.conv.forward_cat([x2, ops.upsample2x(x1, align_corners=True)])
        return self.conv(ops.upcat(x2, x1, align_corners=True))


class BaseDecodeHead(BaseModule):
    """decode_head.py:15-296 for the single-input case used by UnetHead (input_transform=None)."""

    def __init__(self, num_classes=2, in_channels=64, channels=64, *, out_channels=None, threshold=None, dropout_ratio=0.1,
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), in_index=-1, input_transform=None,
                 loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0), ignore_index=255,
                 sampler=None, align_corners=False,
                 init_cfg=dict(type="Normal", std=0.01, override=dict(name="conv_seg"))):
        super().__init__(init_cfg)
        if input_transform is not None:
            raise NotImplementedError("input_transform is not used by UnetHead")
        if not isinstance(in_channels, int) or not isinstance(in_index, int):
            raise AssertionError("in_channels and in_index must be int when input_transform is None")
        self.input_transform = input_transform
        self.in_channels, self.in_index = in_channels, in_index
        self.channels = channels
        self.dropout_ratio = dropout_ratio
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.ignore_index = ignore_index
        self.align_corners = align_corners
        if out_channels is None:
            if num_classes == 2:
                warnings.warn("For binary segmentation, we suggest using `out_channels = 1` to define the output "
                              "channels of segmentor, and use `threshold` to convert seg_logist into a prediction "
                              "applying a threshold")
            out_channels = num_classes
        if out_channels != num_classes and out_channels != 1:
            raise ValueError("out_channels should be equal to num_classes, except binary segmentation set out_channels == 1 "
                             f"and num_classes == 2, but got out_channels={out_channels} and num_classes={num_classes}")
        if out_channels == 1 and threshold is None:
            threshold = 0.3
            warnings.warn("threshold is not defined for binary, and defaults to 0.3")
        self.num_classes, self.out_channels, self.threshold = num_classes, out_channels, threshold
        if isinstance(loss_decode, dict):
            self.loss_decode = build_loss(loss_decode)
        elif isinstance(loss_decode, (list, tuple)):
            self.loss_decode = nn.ModuleList([build_loss(l) for l in loss_decode])
        else:
            raise TypeError(f"loss_decode must be a dict or sequence of dict, but got {type(loss_decode)}")
        if sampler is not None:
            raise NotImplementedError("pixel samplers (OHEM) are outside the STC-UNet path")
        self.sampler = None
        self.conv_seg = nn.Conv2d(channels, self.out_channels, kernel_size=1)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.fp16_enabled = False

    def extra_repr(self):
        return f"input_transform={self.input_transform}, ignore_index={self.ignore_index}, align_corners={self.align_corners}"

    def _dropout_mask(self, feat_nhwc):
        """Dropout2d(p): one Bernoulli(1-p)/(1-p) factor per (sample, channel); sampled with torch's RNG."""
        if self.dropout is None or not self.training:
            return None
        N, C = feat_nhwc.shape[0], feat_nhwc.shape[-1]
        p = self.dropout_ratio
        keep = torch.rand((N, C), device=feat_nhwc.device) >= p
        return keep.float() / (1.0 - p)

    def cls_seg(self, feat_nhwc):
        return ops.cls_seg(feat_nhwc, self.conv_seg, self._dropout_mask(feat_nhwc))

    def forward_train(self, inputs, img_metas, gt_semantic_seg, train_cfg):
        return self.losses(self(inputs), gt_semantic_seg)

    def forward_test(self, inputs, img_metas, test_cfg):
        return self.forward(inputs)

    def losses(self, seg_logit, seg_label):
        """decode_head.py:261-296.  The CE / Dice / accuracy triple is ONE fused kernel pass over the logits;
        each configured loss module picks its component from it (so loss names and weights follow the config)."""
        loss = dict()
        if seg_logit.shape[2:] != seg_label.shape[2:]:
            raise NotImplementedError("seg_logit/seg_label size mismatch: UnetHead always predicts at label resolution")
        seg_label = seg_label.squeeze(1)
        losses_decode = self.loss_decode if isinstance(self.loss_decode, nn.ModuleList) else [self.loss_decode]
        ce, dice, acc = ops.seg_loss(seg_logit.float(), ops.labels_to_int64(seg_label), self.ignore_index, 1.0)
        fused = {"ce": ce, "dice": dice}
        for ld in losses_decode:
            if hasattr(ld, "from_fused"):
                val = ld.from_fused(fused, ignore_index=self.ignore_index)
            else:  # a stock mmseg loss object built from the same config: same maths, its loss_weight
                kind = type(ld).__name__
                if kind not in ("CrossEntropyLoss", "DiceLoss"):
                    raise NotImplementedError(f"loss {kind} is not on the STC-UNet path")
                _check_stock_loss(ld, kind, self.ignore_index)
                val = ops.scale(fused["ce" if kind == "CrossEntropyLoss" else "dice"], getattr(ld, "loss_weight", 1.0))
            loss[ld.loss_name] = val if ld.loss_name not in loss else loss[ld.loss_name] + val
        loss["acc_seg"] = acc.detach()
        return loss


def _check_stock_loss(ld, kind: str, ignore_index: int):
    """A stock mmseg CrossEntropyLoss / DiceLoss object is only honoured when it is configured for exactly what the fused kernel
    computes (cross_entropy_loss.py:186-297, dice_loss.py:50-137 defaults); any other setting must raise, never be silently ignored."""
    bad = []
    if getattr(ld, "class_weight", None) is not None:
        bad.append("class_weight")
    if getattr(ld, "reduction", "mean") != "mean":
        bad.append("reduction")
    if kind == "CrossEntropyLoss":
        bad += [k for k in ("use_sigmoid", "use_mask", "avg_non_ignore") if getattr(ld, k, False)]
    else:
        if getattr(ld, "smooth", 1) != 1:
            bad.append("smooth")
        if getattr(ld, "exponent", 2) != 2:
            bad.append("exponent")
        if getattr(ld, "ignore_index", ignore_index) != ignore_index:
            bad.append("ignore_index")
    if bad:
        raise NotImplementedError(f"{kind} with non-default {', '.join(bad)} is not computed by the fused CE/Dice kernel of stc_unet_b200")


@HEADS.register_module()
class UnetHead(BaseDecodeHead):
    def __init__(self, decoder_channel=[1024, 512, 256, 128, 64], se=False, **kwargs):
        super().__init__(**kwargs)
        dc = decoder_channel
        self.up1 = Up(dc[0], int(dc[0] / 4), se=se)
        self.up2 = Up(dc[1], int(dc[1] / 4), se=se)
        self.up3 = Up(dc[2], int(dc[2] / 4), se=se)
        self.up4 = Up(dc[3], dc[4], se=se)

    def _features(self, inputs):
        f = [ops._chk(_to_nhwc(t)) for t in inputs]
        out = self.up1(f[4], f[3])
        out = self.up2(out, f[2])
        out = self.up3(out, f[1])
        return self.up4(out, f[0])

    def forward(self, inputs):
        return self.cls_seg(self._features(inputs))


# ------------------------------------------------------------------------------------------
# losses (thin views over the fused kernel's outputs)
# ------------------------------------------------------------------------------------------
@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    """cross_entropy_loss.py:186-297 for use_sigmoid=False, use_mask=False, class_weight=None,
    reduction='mean', avg_non_ignore=False (the only configuration on the STC-UNet path)."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction="mean", class_weight=None, loss_weight=1.0,
                 loss_name="loss_ce", avg_non_ignore=False):
        super().__init__()
        if use_sigmoid or use_mask or class_weight is not None or reduction != "mean" or avg_non_ignore:
            raise NotImplementedError("CrossEntropyLoss (stc_unet_b200): only softmax CE, mean over all pixels, no class weights")
        self.use_sigmoid, self.use_mask, self.reduction = use_sigmoid, use_mask, reduction
        self.loss_weight, self.class_weight, self.avg_non_ignore = loss_weight, class_weight, avg_non_ignore
        if not self.avg_non_ignore and self.reduction == "mean":
            warnings.warn("Default ``avg_non_ignore`` is False, if you would like to ignore the certain label and average loss "
                          "over non-ignore labels, which is the same with PyTorch official cross_entropy, set "
                          "``avg_non_ignore=True``.")
        self._loss_name = loss_name

    def from_fused(self, fused, ignore_index=255):
        return ops.scale(fused["ce"], self.loss_weight)

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=-100, **kwargs):
        if weight is not None or avg_factor is not None or reduction_override not in (None, "mean"):
            raise NotImplementedError("CrossEntropyLoss (stc_unet_b200): weight/avg_factor/reduction_override unsupported")
        ce, _, _ = ops.seg_loss(cls_score.float(), ops.labels_to_int64(label), ignore_index, 1.0)
        return ops.scale(ce, self.loss_weight)

    @property
    def loss_name(self):
        return self._loss_name


@LOSSES.register_module()
class DiceLoss(nn.Module):
    """dice_loss.py:50-137 with smooth=1, exponent=2, reduction='mean', class_weight=None."""

    def __init__(self, smooth=1, exponent=2, reduction="mean", class_weight=None, loss_weight=1.0, ignore_index=255,
                 loss_name="loss_dice", **kwargs):
        super().__init__()
        if exponent != 2 or reduction != "mean" or class_weight is not None or smooth != 1:
            raise NotImplementedError("DiceLoss (stc_unet_b200): only smooth=1, exponent=2, reduction='mean', no class weights")
        self.smooth, self.exponent, self.reduction = smooth, exponent, reduction
        self.class_weight, self.loss_weight, self.ignore_index = class_weight, loss_weight, ignore_index
        self._loss_name = loss_name

    def from_fused(self, fused, ignore_index=255):
        if ignore_index != self.ignore_index:
            raise NotImplementedError("DiceLoss ignore_index differs from the head's")
        return ops.scale(fused["dice"], self.loss_weight)

    def forward(self, pred, target, avg_factor=None, reduction_override=None, **kwargs):
        if avg_factor is not None or reduction_override not in (None, "mean"):
            raise NotImplementedError("DiceLoss (stc_unet_b200): avg_factor/reduction_override unsupported")
        _, dice, _ = ops.seg_loss(pred.float(), ops.labels_to_int64(target), self.ignore_index, float(self.smooth))
        return ops.scale(dice, self.loss_weight)

    @property
    def loss_name(self):
        return self._loss_name
