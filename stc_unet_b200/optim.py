"""Optimizer / LR-schedule / checkpoint plumbing on the reference's own interfaces.

  * AdamB200      — `optimizer = dict(type='AdamB200', lr=1e-5, betas=(0.9, 0.999))` in place of my_config/STC-UNet.py:87.  mmcv's
                    DefaultOptimizerConstructor builds it as `cls(params, **cfg)` (mmseg/core/builder.py:22-33, called from
                    mmseg/apis/train.py:121); it is a torch.optim.Optimizer, so OptimizerHook (`zero_grad(); loss.backward();
                    step()`) and the LR hooks (which write `param_groups[i]['lr']`) work unchanged.  All parameters live in ONE flat
                    fp32 buffer and all gradients in ONE flat arena our kernels write into, so `step()` is one stc_adam_step launch.
  * poly_lr / PolyLrUpdater — `lr_config = dict(policy='poly', power=0.9, min_lr=1e-6, by_epoch=True)` (my_config/STC-UNet.py:90):
                    mmcv 1.7's PolyLrUpdaterHook.get_lr restated: (base - min_lr) * (1 - progress / max_progress) ** power + min_lr.
  * save_checkpoint / load_checkpoint — the mmcv checkpoint layout `{'meta': ..., 'state_dict': ..., 'optimizer': ...}` that
                    tools/test.py / mmseg/apis/inference.py:34 read (`load_checkpoint(model, path, map_location='cpu')`): the
                    reference's keys and fp32 tensors only; packed bf16 operand copies are derived caches and never serialised.
"""
from __future__ import annotations

import time
from collections import OrderedDict
from typing import Optional

import torch

from . import ops
from ._lib import lib, stream_ptr
from .registry import HAVE_MMSEG, _LocalRegistry
from .train import FlatParams


class AdamB200(torch.optim.Optimizer):
    """torch.optim.Adam (no amsgrad) with one fused update over flat parameter / gradient / moment buffers."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError("invalid Adam hyper-parameter")          # torch.optim.Adam's own checks
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("AdamB200 keeps ONE flat buffer: paramwise_cfg / several parameter groups are not supported")
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not ps or not all(p.is_cuda and p.dtype == torch.float32 for p in ps):
            raise RuntimeError("AdamB200 needs fp32 CUDA parameters (move the model to the GPU before building the optimizer)")
        self.flat = FlatParams(ps)                # p.data become views into one buffer
        self.arena = ops.GradArena(self.flat.params)
        self.exp_avg = torch.zeros_like(self.flat.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat.flat)
        self.t = 0
        self.steps = {}                           # per-parameter step counts (they differ only if a parameter ever lacked a gradient)
        ops.set_grad_arena(self.arena)            # from now on our backward kernels write parameter gradients straight into the arena

    def zero_grad(self, set_to_none: bool = True):
        for p in self.flat.params:
            p.grad = None
        self.arena._claimed.clear()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        self.t += 1
        # contiguous runs of parameters that HAVE a gradient and share a step count (torch's Adam skips a parameter without a
        # gradient entirely, so its own step count - the bias correction - lags afterwards); normally ONE run = one launch.
        # A gradient that is not already the arena view (produced by a foreign op, or replaced by gradient clipping) is copied in.
        runs, start, cur_t = [], None, None
        for p in self.flat.params:
            off, n = self.arena.offsets[id(p)]
            if p.grad is None:
                if start is not None:
                    runs.append((start, off, cur_t)); start = None
                continue
            view = self.arena.view(p)
            if p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            t = self.steps.get(id(p), 0) + 1
            self.steps[id(p)] = t
            if start is not None and t != cur_t:
                runs.append((start, off, cur_t)); start = None
            if start is None:
                start, cur_t = off, t
        if start is not None:
            runs.append((start, self.arena.total, cur_t))
        for a, b, t in runs:
            lib.call("stc_adam_step", self.flat.flat[a:], self.arena.flat[a:], self.exp_avg[a:], self.exp_avg_sq[a:], b - a, float(g["lr"]),
                     float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), t, stream_ptr())
        ops.invalidate_weight_caches()
        return loss

    def state_dict(self):
        sd = super().state_dict()
        sd["state"] = dict(step=self.t, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq,
                           param_steps=[self.steps.get(id(p), 0) for p in self.flat.params])
        return sd

    def load_state_dict(self, sd):
        st = sd["state"]
        self.t = int(st["step"])
        self.exp_avg.copy_(st["exp_avg"]); self.exp_avg_sq.copy_(st["exp_avg_sq"])
        self.steps = {id(p): int(t) for p, t in zip(self.flat.params, st.get("param_steps", [self.t] * len(self.flat.params)))}
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v


if HAVE_MMSEG:  # pragma: no cover - mmcv is not in the build image
    from mmcv.runner.optimizer import OPTIMIZERS
    OPTIMIZERS.register_module(module=AdamB200, force=True)
else:
    OPTIMIZERS = _LocalRegistry("optimizer")
    OPTIMIZERS.register_module(module=AdamB200)
    OPTIMIZERS.register_module(name="Adam", module=torch.optim.Adam)


def build_optimizer(model, cfg):
    """mmseg.core.build_optimizer for the default constructor without paramwise_cfg (mmseg/core/builder.py:22-33)."""
    cfg = dict(cfg)
    if cfg.pop("paramwise_cfg", None):
        raise NotImplementedError("paramwise_cfg is not supported")
    cfg.pop("constructor", None)
    m = model.module if hasattr(model, "module") else model
    typ = cfg.pop("type")
    cls = OPTIMIZERS.get(typ)
    if cls is None:
        raise KeyError(f"{typ} is not in the optimizer registry")
    return cls([p for p in m.parameters() if p.requires_grad], **cfg)


def poly_lr(base_lr: float, progress: int, max_progress: int, power: float = 1.0, min_lr: float = 0.0) -> float:
    coeff = (1 - progress / max_progress) ** power
    return (base_lr - min_lr) * coeff + min_lr


class PolyLrUpdater:
    """lr_config policy='poly': call before_epoch(epoch) (by_epoch=True, the reference's setting) or before_iter(it)."""

    def __init__(self, optimizer, max_progress: int, power: float = 1.0, min_lr: float = 0.0, by_epoch: bool = True):
        self.opt, self.max_progress, self.power, self.min_lr, self.by_epoch = optimizer, max_progress, power, min_lr, by_epoch
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", g["lr"])

    def _set(self, progress: int):
        for g in self.opt.param_groups:
            g["lr"] = poly_lr(g["initial_lr"], progress, self.max_progress, self.power, self.min_lr)

    def before_epoch(self, epoch: int):
        if self.by_epoch:
            self._set(epoch)

    def before_iter(self, it: int):
        if not self.by_epoch:
            self._set(it)


def _plain_state_dict(model) -> "OrderedDict[str, torch.Tensor]":
    m = model.module if hasattr(model, "module") else model
    return OrderedDict((k, v.detach().cpu().clone()) for k, v in m.state_dict().items())


def save_checkpoint(model, filename: str, optimizer=None, meta: Optional[dict] = None):
    meta = dict(meta or {})
    meta.setdefault("time", time.asctime())
    m = model.module if hasattr(model, "module") else model
    if hasattr(m, "CLASSES") and m.CLASSES is not None:
        meta.setdefault("CLASSES", m.CLASSES)
    ckpt = dict(meta=meta, state_dict=_plain_state_dict(model))
    if optimizer is not None:
        osd = optimizer.state_dict()
        ckpt["optimizer"] = {k: ({kk: (vv.detach().cpu() if torch.is_tensor(vv) else vv) for kk, vv in v.items()} if isinstance(v, dict) else v)
                             for k, v in osd.items()}
    torch.save(ckpt, filename)


def load_checkpoint(model, filename: str, map_location="cpu", strict: bool = False, revise_keys=((r"^module\.", ""),)):
    """Returns the checkpoint dict.  Accepts `{'state_dict': ...}` (mmcv) or a bare state_dict; parameters are copied IN PLACE, so a
    model whose parameters already are views into an optimizer's flat buffer keeps them."""
    import re
    ckpt = torch.load(filename, map_location=map_location, weights_only=False)
    if not isinstance(ckpt, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {filename}")
    sd = ckpt.get("state_dict", ckpt)
    for pat, rep in revise_keys:
        sd = OrderedDict((re.sub(pat, rep, k), v) for k, v in sd.items())
    m = model.module if hasattr(model, "module") else model
    own = m.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    mismatched = [k for k in sd if k in own and tuple(own[k].shape) != tuple(sd[k].shape)]
    if strict and (missing or unexpected or mismatched):
        raise RuntimeError(f"load_checkpoint: missing {missing}, unexpected {unexpected}, size mismatch {mismatched}")
    with torch.no_grad():
        for k, v in sd.items():
            if k in own and k not in mismatched:
                own[k].copy_(v)
    ops.invalidate_weight_caches()
    return ckpt
