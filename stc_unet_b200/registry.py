"""Registry boundary (mmseg/models/builder.py:8-49).

When real mmcv + mmseg are importable the modules register into mmseg's own `MODELS` registry
(`BACKBONES = HEADS = LOSSES = SEGMENTORS = MODELS`) under `<Name>B200` and — with
`register_module(force=True)` when `STC_B200_OVERRIDE=1` — over the original names, so
`my_config/STC-UNet.py`, `tools/train.py` and `tools/test.py` run unchanged
(`custom_imports=dict(imports=['stc_unet_b200'])`).  Offline (this image has neither mmcv nor
mmseg) a minimal registry with the same `register_module()` / `build(cfg)` behaviour is used.
"""
from __future__ import annotations

import importlib.util
import os

import torch.nn as nn

def _have(name: str) -> bool:
    import sys
    mod = sys.modules.get(name)
    if mod is not None:   # already imported: real package unless it is the oracle's import stub
        return not getattr(sys.modules.get("mmcv"), "_stc_stub", False) and getattr(mod, "__spec__", None) is not None
    try:
        return importlib.util.find_spec(name) is not None
    except (ValueError, ImportError):
        return False


HAVE_MMSEG = _have("mmcv") and _have("mmseg")


class _LocalRegistry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force and self.module_dict[key] is not cls:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise KeyError('`cfg` must be a dict containing the key "type"')
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**args)


class _LocalBaseModule(nn.Module):
    """mmcv.runner.BaseModule subset: init_cfg + init_weights() (Normal/override only, which is all
    decode_head.py:78-79 uses; UnetBackbone has init_cfg=None -> torch default init)."""

    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        cfg = self.init_cfg
        if isinstance(cfg, dict) and cfg.get("type") == "Normal":
            ov = cfg.get("override")
            if ov and hasattr(self, ov["name"]):
                m = getattr(self, ov["name"])
                nn.init.normal_(m.weight, mean=cfg.get("mean", 0.0), std=cfg.get("std", 0.01))
                if getattr(m, "bias", None) is not None:
                    nn.init.constant_(m.bias, cfg.get("bias", 0.0))
        for c in self.children():
            if hasattr(c, "init_weights"):
                c.init_weights()


if HAVE_MMSEG:  # pragma: no cover - not available in the build image
    from mmcv.runner import BaseModule as _MMBase
    from mmseg.models.builder import MODELS as _MM

    class _Proxy:
        """Registers `<Name>B200` always, and the bare name too when STC_B200_OVERRIDE=1."""

        def __init__(self, reg):
            self.reg = reg

        def register_module(self, name=None, force=False, module=None):
            def deco(cls):
                base = name or cls.__name__
                self.reg.register_module(name=base + "B200", force=True, module=cls)
                if os.environ.get("STC_B200_OVERRIDE", "0") == "1":
                    self.reg.register_module(name=base, force=True, module=cls)
                return cls
            return deco(module) if module is not None else deco

        def build(self, cfg, default_args=None):
            return self.reg.build(cfg, default_args=default_args)

        def get(self, key):
            return self.reg.get(key)

    MODELS = _Proxy(_MM)
    BaseModule = _MMBase
else:
    MODELS = _LocalRegistry("models")
    BaseModule = _LocalBaseModule

BACKBONES = NECKS = HEADS = LOSSES = SEGMENTORS = MODELS


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    return SEGMENTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))
