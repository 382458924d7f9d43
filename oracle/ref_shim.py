"""ORACLE / TEST INFRASTRUCTURE ONLY — never imported by the product path.

Loads the reference's *own* hot-path source files, unchanged, from
``/root/reference`` (present only in the build container, never on the GPU
box) under a minimal stub of the mmcv / timm symbols they touch.  Used to
(1) pin ``oracle/stc_oracle.py`` (the restatement that travels) and
(2) generate the golden fixtures under ``tests/golden/``.

Files executed (reference paths):
  mmseg/ops/wrappers.py, mmseg/models/losses/{utils,accuracy,cross_entropy_loss,dice_loss}.py,
  mmseg/models/decode_heads/{decode_head,unet_head}.py,
  mmseg/models/backbones/unet_backbone.py, mmseg/core/evaluation/metrics.py

The stubbed symbols follow SURVEY.md Appendix C.  Nothing here restates
reference arithmetic: the arithmetic is the reference's own code.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("STC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "mmseg/models/backbones/unet_backbone.py"))


class _Registry:
    """Tiny stand-in for mmcv.utils.Registry (register_module decorator + build)."""

    def __init__(self, name="models", parent=None, **_):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        cfg = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        typ = cfg.pop("type")
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        return cls(**cfg)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        # mmcv applies init_cfg; the only one on this path is
        # Normal(std=0.01, override=conv_seg) (decode_head.py:78-79).
        cfg = self.init_cfg
        if isinstance(cfg, dict) and cfg.get("type") == "Normal":
            ov = cfg.get("override")
            if ov and hasattr(self, ov["name"]):
                m = getattr(self, ov["name"])
                nn.init.normal_(m.weight, mean=0.0, std=cfg.get("std", 0.01))
                if getattr(m, "bias", None) is not None:
                    nn.init.constant_(m.bias, 0.0)
        for c in self.children():
            if hasattr(c, "init_weights"):
                c.init_weights()


def _identity_decorator_factory(*_a, **_k):
    def deco(fn):
        return fn
    return deco


class _RefConvModule(nn.Module):
    """Restatement of mmcv.cnn.ConvModule (mmcv-full 1.7.0, not in this image; semantics from its documented behaviour:
    order conv -> norm -> act, bias='auto' => no conv bias when a norm layer follows, children named conv / bn / activate,
    ReLU inplace).  Call sites: unet.py:66-75,199-208; up_conv_block.py:84-93; fcn_head.py:45-65.  Not pinned by any numeric
    test of the reference."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias="auto",
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True, **_):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                              groups=groups, bias=bias)
        if self.with_norm:
            self.bn = nn.SyncBatchNorm(out_channels) if norm_cfg.get("type") == "SyncBN" else nn.BatchNorm2d(out_channels)
        if self.with_activation:
            assert act_cfg.get("type") == "ReLU"
            self.activate = nn.ReLU(inplace=inplace)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def _mod(name):
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


_LOADED = {}


def _load(dotted, relpath):
    if dotted in _LOADED:
        return _LOADED[dotted]
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(dotted, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    _LOADED[dotted] = mod
    return mod


def load_reference():
    """Returns a namespace with the reference's classes/functions for this path."""
    if "ns" in _LOADED:
        return _LOADED["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    if "mmcv" in sys.modules and not getattr(sys.modules["mmcv"], "_stc_stub", False):
        raise RuntimeError("a real mmcv is imported; the shim must not shadow it")

    R = _Registry("models")
    mmcv = _mod("mmcv")
    mmcv._stc_stub = True
    mmcv.load = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("mmcv.load stub"))
    mmcv.imread = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("mmcv.imread stub"))
    mmcv.is_list_of = lambda seq, t: isinstance(seq, list) and all(isinstance(s, t) for s in seq)
    cnn = _mod("mmcv.cnn")
    cnn.MODELS = R
    cnn.ConvModule = _RefConvModule
    cnn.UPSAMPLE_LAYERS = _Registry("upsample")
    cnn.build_activation_layer = lambda cfg: nn.ReLU(inplace=cfg.get("inplace", False))
    cnn.build_norm_layer = lambda cfg, n, postfix="": ("bn" + str(postfix), nn.BatchNorm2d(n))

    def _build_upsample_layer(cfg, *args, **kwargs):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        if typ == "deconv":
            return nn.ConvTranspose2d(*args, **kwargs, **cfg)
        return cnn.UPSAMPLE_LAYERS.module_dict[typ](*args, **kwargs, **cfg)
    cnn.build_upsample_layer = _build_upsample_layer
    bricks = _mod("mmcv.cnn.bricks")
    reg = _mod("mmcv.cnn.bricks.registry")
    reg.NORM_LAYERS = _Registry("norm")
    reg.ATTENTION = _Registry("attention")
    runner = _mod("mmcv.runner")
    runner.BaseModule = _BaseModule
    runner.auto_fp16 = _identity_decorator_factory
    runner.force_fp32 = _identity_decorator_factory
    utils = _mod("mmcv.utils")
    utils.Registry = _Registry
    pw = _mod("mmcv.utils.parrots_wrapper")
    pw.SyncBatchNorm = nn.SyncBatchNorm
    pw._BatchNorm = nn.modules.batchnorm._BatchNorm
    mmcv.cnn, mmcv.runner, mmcv.utils = cnn, runner, utils
    cnn.bricks = bricks
    bricks.registry = reg
    utils.parrots_wrapper = pw

    timm = _mod("timm")
    tm = _mod("timm.models")
    tl = _mod("timm.models.layers")
    tl.DropPath = nn.Identity
    tl.to_2tuple = lambda x: (x, x)
    tl.trunc_normal_ = nn.init.trunc_normal_
    timm.models, tm.layers = tm, tl

    mmseg = _mod("mmseg")
    core = _mod("mmseg.core")
    core.build_pixel_sampler = lambda cfg, **k: None
    core.add_prefix = lambda d, p: {f"{p}.{k}": v for k, v in d.items()}
    ops = _mod("mmseg.ops")
    models = _mod("mmseg.models")
    builder = _mod("mmseg.models.builder")
    for n in ("MODELS", "BACKBONES", "NECKS", "HEADS", "LOSSES", "SEGMENTORS"):
        setattr(builder, n, R)
    builder.ATTENTION = reg.ATTENTION
    builder.build_loss = R.build
    builder.build_backbone = R.build
    builder.build_head = R.build
    _mod("mmseg.models.backbones")
    _mod("mmseg.models.decode_heads")
    losses = _mod("mmseg.models.losses")
    mmseg.core, mmseg.ops, mmseg.models = core, ops, models
    models.builder = builder

    w = _load("mmseg.ops.wrappers", "mmseg/ops/wrappers.py")
    ops.resize, ops.Upsample = w.resize, w.Upsample
    lu = _load("mmseg.models.losses.utils", "mmseg/models/losses/utils.py")
    acc = _load("mmseg.models.losses.accuracy", "mmseg/models/losses/accuracy.py")
    losses.accuracy, losses.Accuracy = acc.accuracy, acc.Accuracy
    losses.utils = lu
    ce = _load("mmseg.models.losses.cross_entropy_loss", "mmseg/models/losses/cross_entropy_loss.py")
    dl = _load("mmseg.models.losses.dice_loss", "mmseg/models/losses/dice_loss.py")
    dh = _load("mmseg.models.decode_heads.decode_head", "mmseg/models/decode_heads/decode_head.py")
    uh = _load("mmseg.models.decode_heads.unet_head", "mmseg/models/decode_heads/unet_head.py")
    ub = _load("mmseg.models.backbones.unet_backbone", "mmseg/models/backbones/unet_backbone.py")
    mutils = _mod("mmseg.models.utils")
    ucb = _load("mmseg.models.utils.up_conv_block", "mmseg/models/utils/up_conv_block.py")
    mutils.UpConvBlock = ucb.UpConvBlock
    unet_b = _load("mmseg.models.backbones.unet", "mmseg/models/backbones/unet.py")
    fcn = _load("mmseg.models.decode_heads.fcn_head", "mmseg/models/decode_heads/fcn_head.py")
    _mod("mmseg.core.evaluation")
    met = _load("mmseg.core.evaluation.metrics", "mmseg/core/evaluation/metrics.py")

    ns = types.SimpleNamespace(
        registry=R, UnetBackbone=ub.UnetBackbone, UnetHead=uh.UnetHead,
        BaseDecodeHead=dh.BaseDecodeHead, CrossEntropyLoss=ce.CrossEntropyLoss,
        DiceLoss=dl.DiceLoss, accuracy=acc.accuracy, resize=w.resize,
        intersect_and_union=met.intersect_and_union, backbone_mod=ub, head_mod=uh,
        KernelSelectAttention=ub.KernelSelectAttention, TransformerBlock=ub.TransformerBlock,
        CoordAtt=uh.CoordAtt, metrics_mod=met, UNet=unet_b.UNet, FCNHead=fcn.FCNHead, BasicConvBlock=unet_b.BasicConvBlock,
        InterpConv=unet_b.InterpConv, UpConvBlock=ucb.UpConvBlock)
    _LOADED["ns"] = ns
    return ns


def revert_sync_batchnorm(module: nn.Module) -> nn.Module:
    """SyncBN -> BN keeping parameters/buffers/keys (what tools/train.py:205-210
    does through mmcv for non-distributed runs)."""
    out = module
    if isinstance(module, nn.SyncBatchNorm):
        out = nn.BatchNorm2d(module.num_features, module.eps, module.momentum,
                             module.affine, module.track_running_stats)
        if module.affine:
            out.weight, out.bias = module.weight, module.bias
        out.running_mean, out.running_var = module.running_mean, module.running_var
        out.num_batches_tracked = module.num_batches_tracked
        out.training = module.training
    for name, child in module.named_children():
        out.add_module(name, revert_sync_batchnorm(child))
    return out


LOSS_CFG = [
    dict(type="CrossEntropyLoss", use_sigmoid=False, loss_name="loss_bce", loss_weight=1.0),
    dict(type="DiceLoss", loss_name="loss_dice", loss_weight=1.0),
]


def build_reference_model(stc: bool, num_classes: int, dropout_ratio: float = 0.0, seed: int = 0):
    """(backbone, head) exactly as my_config/STC-UNet.py:3-20 (stc=True) or
    my_config/U-Net.py:3-17 (stc=False) build them, SyncBN reverted."""
    ns = load_reference()
    torch.manual_seed(seed)
    if stc:
        bb = ns.UnetBackbone(in_channels=3, context_layer="kernelselect", transformer_block=True,
                             channel_list=[64, 128, 256, 512])
        hd = ns.UnetHead(se=True, num_classes=num_classes, channels=64, threshold=0.2,
                         norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG,
                         dropout_ratio=dropout_ratio)
    else:
        bb = ns.UnetBackbone(in_channels=3, channel_list=[64, 128, 256, 512])
        hd = ns.UnetHead(num_classes=num_classes, channels=64, threshold=0.2,
                         norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG,
                         dropout_ratio=dropout_ratio)
    bb.init_weights()
    hd.init_weights()
    return revert_sync_batchnorm(bb), revert_sync_batchnorm(hd)


def build_reference_model_b(num_classes: int, dropout_ratio: float = 0.0, seed: int = 0, base_channels: int = 64, num_stages: int = 5,
                            upsample: str = "InterpConv"):
    """(UNet, FCNHead) as configs/_base_/models/fcn_unet_s5-d16.py:3-34 builds them (BN instead of SyncBN, torch default init:
    mmcv's Kaiming init_cfg is not restated — parity tests share the state_dict instead)."""
    ns = load_reference()
    torch.manual_seed(seed)
    norm_cfg = dict(type="BN", requires_grad=True)
    bb = ns.UNet(in_channels=3, base_channels=base_channels, num_stages=num_stages, strides=(1,) * num_stages,
                 enc_num_convs=(2,) * num_stages, dec_num_convs=(2,) * (num_stages - 1), downsamples=(True,) * (num_stages - 1),
                 enc_dilations=(1,) * num_stages, dec_dilations=(1,) * (num_stages - 1), with_cp=False, conv_cfg=None,
                 norm_cfg=norm_cfg, act_cfg=dict(type="ReLU"), upsample_cfg=dict(type=upsample), norm_eval=False)
    hd = ns.FCNHead(in_channels=base_channels, in_index=num_stages - 1, channels=base_channels, num_convs=1, concat_input=False,
                    dropout_ratio=dropout_ratio, num_classes=num_classes, norm_cfg=norm_cfg, align_corners=False,
                    loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0))
    return bb, hd
