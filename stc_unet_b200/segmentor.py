"""EncoderDecoder mirror (mmseg/models/segmentors/encoder_decoder.py:13-280, base.py:112-215) for
environments without mmseg, plus the B200-native replacements of its post-processing:

  * slide_inference: all windows of the batch are stacked into ONE forward (legal in eval mode: BN uses
    running statistics), then accumulated with stc_slide_accum; the softmax is skipped because argmax of
    softmax(preds / count) == argmax(preds / count) (stc_argmax keeps the first maximum like torch.argmax).
  * simple_test returns the reference's `list[np.ndarray int64 (H, W)]`, and `simple_test_device` keeps the
    prediction on the GPU for the on-device confusion-matrix histogram.
  * train_step/_parse_losses: one packed all-reduce and ONE device->host copy for all log vars
    (the reference does an all-reduce + .item() per variable, base.py:196-213).

With real mmseg installed its own EncoderDecoder works unchanged with our backbone/head (they are plain
nn.Modules honouring the registry contract); this class is what bench/tests drive offline.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .registry import SEGMENTORS, BaseModule, build_backbone, build_head


def add_prefix(d, prefix):
    return {f"{prefix}.{k}": v for k, v in d.items()}


def slide_windows(h_img, w_img, crop, stride):
    """Window boxes exactly as encoder_decoder.py:164-179 enumerates them."""
    (hc, wc), (hs, ws) = crop, stride
    hg = max(h_img - hc + hs - 1, 0) // hs + 1
    wg = max(w_img - wc + ws - 1, 0) // ws + 1
    out = []
    for i in range(hg):
        for j in range(wg):
            y2, x2 = min(i * hs + hc, h_img), min(j * ws + wc, w_img)
            out.append((max(y2 - hc, 0), max(x2 - wc, 0), y2, x2))
    return out


@SEGMENTORS.register_module()
class EncoderDecoder(BaseModule):
    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None, init_cfg=None):
        super().__init__(init_cfg)
        if neck is not None or auxiliary_head is not None:
            raise NotImplementedError("neck / auxiliary_head are not part of the STC-UNet configs")
        self.backbone = build_backbone(backbone) if isinstance(backbone, dict) else backbone
        self.decode_head = build_head(decode_head) if isinstance(decode_head, dict) else decode_head
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.out_channels = self.decode_head.out_channels
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg or dict(mode="whole")

    # ---------------------------------------------------------------- training
    def extract_feat(self, img):
        return self.backbone(img)

    def forward_train(self, img, img_metas, gt_semantic_seg):
        x = self.extract_feat(img)
        losses = dict()
        losses.update(add_prefix(self.decode_head.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg), "decode"))
        return losses

    def forward(self, img, img_metas=None, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.simple_test(img, img_metas, **kwargs)

    @staticmethod
    def _parse_losses(losses):
        """Sum of every entry whose name contains 'loss' (base.py:170-215).  Log vars are packed into one
        tensor: one all-reduce, one D2H copy, done lazily by `log_vars_to_host`."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value
            elif isinstance(value, list):
                raise NotImplementedError("list-valued losses are not produced on this path")
            else:
                raise TypeError(f"{name} is not a tensor or list of tensors")
        terms = [v for k, v in log_vars.items() if "loss" in k]
        loss = terms[0]
        for t in terms[1:]:
            loss = ops.add_autograd(loss.reshape(1), t.reshape(1)).reshape(())
        log_vars["loss"] = loss
        return loss, log_vars

    @staticmethod
    def log_vars_to_host(log_vars):
        names = list(log_vars)
        packed = torch.stack([log_vars[n].detach().float().reshape(()) for n in names])
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(packed)
            packed = packed / dist.get_world_size()
        vals = packed.cpu().tolist()
        return OrderedDict(zip(names, vals))

    def train_step(self, data_batch, optimizer=None, **kwargs):
        losses = self(**data_batch)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data_batch["img"]))

    # ---------------------------------------------------------------- inference
    def encode_decode(self, img, img_metas=None):
        return self.decode_head.forward_test(self.extract_feat(img), img_metas, self.test_cfg)

    def _slide_logits(self, img):
        h_crop, w_crop = self.test_cfg["crop_size"]
        u8 = img.dtype == torch.uint8          # decoded (N, H, W, C) pixels: normalised on the device by the backbone (ops.image_to_nhwc)
        N, H, W = (img.shape[0], img.shape[1], img.shape[2]) if u8 else (img.shape[0], img.shape[2], img.shape[3])
        wins = slide_windows(H, W, (h_crop, w_crop), tuple(self.test_cfg["stride"]))
        preds = torch.zeros((N, self.out_channels, H, W), dtype=torch.float32, device=img.device)
        count = torch.zeros((N, H, W), dtype=torch.float32, device=img.device)
        max_windows = int(self.test_cfg.get("max_windows_per_forward", 64))
        group = max(1, max_windows // max(N, 1))
        for g0 in range(0, len(wins), group):
            chunk = wins[g0:g0 + group]
            # windows that share a shape are batched along N (all of them do unless the image is smaller than the crop)
            crops = torch.cat([(img[:, y1:y2, x1:x2] if u8 else img[:, :, y1:y2, x1:x2]) for (y1, x1, y2, x2) in chunk], dim=0)
            logits = self.encode_decode(crops)
            for wi, (y1, x1, y2, x2) in enumerate(chunk):
                ops.slide_accum(logits[wi * N:(wi + 1) * N], preds, count, y1, x1)
        return preds, count

    def inference_device(self, img) -> torch.Tensor:
        """(N,H,W) int64 prediction on the device (whole or slide mode)."""
        mode = self.test_cfg.get("mode", "whole")
        assert mode in ("slide", "whole")
        with torch.no_grad():
            if mode == "slide":
                preds, count = self._slide_logits(img)
                return ops.argmax_nchw(preds, count)
            return ops.argmax_nchw(self.encode_decode(img), None)

    def slide_inference(self, img, img_meta=None, rescale=False):
        with torch.no_grad():
            preds, count = self._slide_logits(img)
        return preds / count.unsqueeze(1)

    def whole_inference(self, img, img_meta=None, rescale=False):
        with torch.no_grad():
            return self.encode_decode(img)

    def simple_test(self, img, img_meta=None, rescale=True) -> List:
        seg_pred = self.inference_device(img).cpu().numpy()
        return list(seg_pred)
