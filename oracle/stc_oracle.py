"""ORACLE / TEST INFRASTRUCTURE ONLY — never imported by the product path.

Plain-PyTorch fp32 restatement of the reference's hot path (STC-UNet / U-Net
encoder-decoder forward, CE+Dice loss, accuracy, slide-window grid, integer
area histograms).  It is written functionally over the reference's
``state_dict`` keys so that the same weights drive the reference modules, this
oracle and the B200 modules.  Device agnostic (CPU is authoritative; tests may
run it on CUDA for the full-size cases).  Gradients come from torch autograd
over these formulas.

Pinned (tests/test_oracle_vs_reference.py, tests/test_golden.py) against
  (a) the reference's own modules imported from /root/reference through
      oracle/ref_shim.py (build container only), and
  (b) the committed fixtures tests/golden/*.pt produced by oracle/make_golden.py
      from those reference modules, and
  (c) the reference's known-answer tests for CE and the confusion matrix
      (tests/test_models/test_losses/test_ce_loss.py:25-86, tests/test_metrics.py:9-26).

Each function cites the reference lines it restates (paths under /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def batch_norm(sd: SD, p: str, x: torch.Tensor, train: bool, new_stats: Optional[dict],
               eps: float = 1e-5, momentum: float = 0.1) -> torch.Tensor:
    """SyncBatchNorm(C) with torch defaults, reverted to BatchNorm when not
    distributed (unet_backbone.py:64,121,124; unet_head.py:68,71,125;
    tools/train.py:205-210).  x is (N,C,*)."""
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    if train:
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=dims)
        var = ((x - mean.view(shape)) ** 2).mean(dim=dims)  # biased
        if new_stats is not None:
            with torch.no_grad():
                new_stats[p + ".running_mean"] = (1 - momentum) * sd[p + ".running_mean"] + momentum * mean
                new_stats[p + ".running_var"] = (1 - momentum) * sd[p + ".running_var"] + \
                    momentum * var * (n / max(n - 1, 1))
                new_stats[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[p + ".running_mean"], sd[p + ".running_var"]
    xh = (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + eps)
    return xh * sd[p + ".weight"].view(shape) + sd[p + ".bias"].view(shape)


def double_conv(sd: SD, p: str, x, train, ns):
    """DoubleConv: [Conv3x3(bias) -> BN -> ReLU] x2 (unet_backbone.py:116-130; unet_head.py:63-77)."""
    for i in (0, 3):
        x = F.conv2d(x, sd[f"{p}.{i}.weight"], sd[f"{p}.{i}.bias"], padding=1)
        x = torch.relu(batch_norm(sd, f"{p}.{i + 1}", x, train, ns))
    return x


def kernel_select_attention(sd: SD, p: str, x, train, ns):
    """KernelSelectAttention (unet_backbone.py:55-99): k in {3,5,7} conv-BN-ReLU branches,
    U = sum, S = GAP(U), Z = fc(S), a_k = fcs_k(Z), softmax over k, V = sum_k w_k f_k."""
    feats = []
    for bi, k in enumerate((3, 5, 7)):
        f = F.conv2d(x, sd[f"{p}.convs.{bi}.0.weight"], sd[f"{p}.convs.{bi}.0.bias"], padding=k // 2)
        feats.append(torch.relu(batch_norm(sd, f"{p}.convs.{bi}.1", f, train, ns)))
    U = feats[0] + feats[1] + feats[2]
    S = U.mean(dim=(2, 3))
    Z = S @ sd[f"{p}.fc.weight"].t() + sd[f"{p}.fc.bias"]
    a = torch.stack([Z @ sd[f"{p}.fcs.{i}.weight"].t() + sd[f"{p}.fcs.{i}.bias"] for i in range(3)], 0)
    w = torch.softmax(a, dim=0)  # (3,N,C)
    return sum(w[i][:, :, None, None] * feats[i] for i in range(3))


def multihead_attention(sd: SD, p: str, q, k, v, heads: int):
    """torch nn.MultiheadAttention(E, heads) forward, seq-first (L,N,E), no masks, dropout 0."""
    L, N, E = q.shape
    hd = E // heads
    W, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    qp = q @ W[:E].t() + b[:E]
    kp = k @ W[E:2 * E].t() + b[E:2 * E]
    vp = v @ W[2 * E:].t() + b[2 * E:]

    def split(t):  # (L,N,E) -> (N*heads, L, hd)
        return t.reshape(L, N * heads, hd).transpose(0, 1)
    qh, kh, vh = split(qp), split(kp), split(vp)
    att = torch.softmax((qh @ kh.transpose(1, 2)) / math.sqrt(hd), dim=-1)
    o = (att @ vh).transpose(0, 1).reshape(L, N, E)
    return o @ sd[p + ".out_proj.weight"].t() + sd[p + ".out_proj.bias"]


def transformer_block(sd: SD, p: str, x, heads: int = 2, layers: int = 4):
    """TransformerBlock(512,512,2,4) (unet_backbone.py:229-246) with TransformerLayer (:195-209):
    t = p + linear(p); per layer t = MHA(q(t),k(t),v(t)) + t ; t = fc2(fc1(t)) + t."""
    N, C, H, W = x.shape
    t = x.flatten(2).permute(2, 0, 1)
    t = t + (t @ sd[p + ".linear.weight"].t() + sd[p + ".linear.bias"])
    for i in range(layers):
        lp = f"{p}.tr.{i}"
        a = multihead_attention(sd, lp + ".ma", t @ sd[lp + ".q.weight"].t(), t @ sd[lp + ".k.weight"].t(),
                                t @ sd[lp + ".v.weight"].t(), heads)
        t = a + t
        t = (t @ sd[lp + ".fc1.weight"].t()) @ sd[lp + ".fc2.weight"].t() + t
    return t.permute(1, 2, 0).reshape(N, C, H, W)


def backbone_forward(sd: SD, x, train: bool = True, new_stats: Optional[dict] = None) -> List[torch.Tensor]:
    """UnetBackbone.forward (unet_backbone.py:36-52).  KSA / transformer are applied iff their
    keys are in the state_dict (context_layer='kernelselect', transformer_block=True)."""
    x1 = double_conv(sd, "inc.conv.conv", x, train, new_stats)
    feats = [x1]
    cur = x1
    for i in range(1, 5):
        cur = double_conv(sd, f"down{i}.down_conv.1.conv", F.max_pool2d(cur, 2), train, new_stats)
        feats.append(cur)
    x1, x2, x3, x4, x5 = feats
    if "context_layer1_1.fc.weight" in sd:
        r1 = kernel_select_attention(sd, "context_layer1_1", x1, train, new_stats)
        r2 = kernel_select_attention(sd, "context_layer2_1", x2, train, new_stats)
        r3 = kernel_select_attention(sd, "context_layer3_1", x3, train, new_stats)
        x1, x2, x3 = x1 + r1, x2 + r2, x3 + r3
    if "aspp4.linear.weight" in sd:
        x4 = transformer_block(sd, "aspp4", x4) + x4
        x5 = transformer_block(sd, "aspp5", x5) + x5
    return [x1, x2, x3, x4, x5]


def coord_att(sd: SD, p: str, x, train, ns):
    """CoordAtt.forward (unet_head.py:131-146): returns a_w * a_h (an additive map, see :57)."""
    n, c, h, w = x.shape
    xh = x.mean(dim=3, keepdim=True)                       # (n,c,h,1)
    xw = x.mean(dim=2, keepdim=True).permute(0, 1, 3, 2)   # (n,c,w,1)
    y = torch.cat([xh, xw], dim=2)
    y = F.conv2d(y, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
    y = batch_norm(sd, p + ".bn1", y, train, ns)
    y = y * torch.clamp(y + 3, 0, 6) / 6                   # h_swish
    yh, yw = y[:, :, :h], y[:, :, h:].permute(0, 1, 3, 2)
    ah = torch.sigmoid(F.conv2d(yh, sd[p + ".conv_h.weight"], sd[p + ".conv_h.bias"]))
    aw = torch.sigmoid(F.conv2d(yw, sd[p + ".conv_w.weight"], sd[p + ".conv_w.bias"]))
    return aw * ah


def up_block(sd: SD, p: str, x1, x2, train, ns):
    """Up.forward (unet_head.py:50-60): bilinear x2 align_corners=True, pad, cat[skip, up],
    optional `ca(x) + x`, DoubleConv."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    x = torch.cat([x2, x1], dim=1)
    if p + ".ca.conv1.weight" in sd:
        x = coord_att(sd, p + ".ca", x, train, ns) + x
    return double_conv(sd, p + ".conv.conv", x, train, ns)


def head_forward(sd: SD, feats: List[torch.Tensor], train: bool = True, new_stats: Optional[dict] = None,
                 dropout_mask: Optional[torch.Tensor] = None):
    """UnetHead.forward (unet_head.py:26-32) + cls_seg (decode_head.py:254-259).
    dropout_mask: optional (N,64,1,1) already-scaled Dropout2d mask; None = dropout_ratio 0 / eval."""
    out = up_block(sd, "up1", feats[4], feats[3], train, new_stats)
    out = up_block(sd, "up2", out, feats[2], train, new_stats)
    out = up_block(sd, "up3", out, feats[1], train, new_stats)
    out = up_block(sd, "up4", out, feats[0], train, new_stats)
    if dropout_mask is not None:
        out = out * dropout_mask
    return F.conv2d(out, sd["conv_seg.weight"], sd["conv_seg.bias"])


# --------------------------------------------------------------------------
# losses (decode_head.py:261-296)
# --------------------------------------------------------------------------
def cross_entropy_loss(logits, label, ignore_index: int = 255, class_weight=None):
    """CrossEntropyLoss(avg_non_ignore=False): per-pixel CE, 0 on ignored pixels, mean over ALL
    pixels (cross_entropy_loss.py:45-61; utils.py:68-69)."""
    lsm = torch.log_softmax(logits.float(), dim=1)
    valid = label != ignore_index
    idx = torch.where(valid, label, torch.zeros_like(label)).long()
    nll = -lsm.gather(1, idx.unsqueeze(1)).squeeze(1)
    if class_weight is not None:
        nll = nll * torch.as_tensor(class_weight, dtype=nll.dtype, device=nll.device)[idx]
    return (nll * valid).sum() / label.numel()


def dice_loss(logits, label, ignore_index: int = 255, smooth: float = 1.0):
    """DiceLoss.forward -> dice_loss -> binary_dice_loss (dice_loss.py:92-123,13-47): exponent 2,
    one-hot of clamp(label), valid mask in the numerator only, mean over images, sum over
    classes / C."""
    N, C = logits.shape[:2]
    p = torch.softmax(logits.float(), dim=1).reshape(N, C, -1)
    lab = label.reshape(N, -1)
    t = F.one_hot(lab.clamp(0, C - 1).long(), C).permute(0, 2, 1).to(p.dtype)
    m = (lab != ignore_index).to(p.dtype).unsqueeze(1)
    num = 2 * (p * t * m).sum(-1) + smooth
    den = (p * p + t * t).sum(-1) + smooth
    per = 1 - num / den                       # (N,C)
    total = 0
    for i in range(C):
        if i != ignore_index:
            total = total + per[:, i].mean()
    return total / C


def accuracy(logits, label, ignore_index: int = 255):
    """accuracy(topk=1) (accuracy.py:6-61)."""
    pred = logits.argmax(dim=1)
    valid = label != ignore_index
    eps = torch.finfo(torch.float32).eps
    correct = ((pred == label) & valid).float().sum() + eps
    return correct * (100.0 / (valid.sum().item() + eps))


def losses(logits, seg_label, ignore_index: int = 255) -> Dict[str, torch.Tensor]:
    """BaseDecodeHead.losses with loss_decode=[CE(loss_bce), Dice(loss_dice)] (my_config/STC-UNet.py:17-19).
    seg_label is (N,1,H,W) int64; logits already at label size (resize is the identity)."""
    lab = seg_label.squeeze(1)
    return {"loss_bce": cross_entropy_loss(logits, lab, ignore_index),
            "loss_dice": dice_loss(logits, lab, ignore_index),
            "acc_seg": accuracy(logits, lab, ignore_index)}


def forward_train(bsd: SD, hsd: SD, img, gt, train: bool = True,
                  new_stats_b: Optional[dict] = None, new_stats_h: Optional[dict] = None):
    feats = backbone_forward(bsd, img, train, new_stats_b)
    logits = head_forward(hsd, feats, train, new_stats_h)
    out = losses(logits, gt)
    out["logits"] = logits
    return out


# --------------------------------------------------------------------------
# inference post-processing (encoder_decoder.py:157-203,227-280)
# --------------------------------------------------------------------------
def slide_windows(h_img: int, w_img: int, crop: Tuple[int, int], stride: Tuple[int, int]):
    """Window origins exactly as slide_inference computes them (encoder_decoder.py:164-179)."""
    (hc, wc), (hs, ws) = crop, stride
    hg = max(h_img - hc + hs - 1, 0) // hs + 1
    wg = max(w_img - wc + ws - 1, 0) // ws + 1
    wins = []
    for i in range(hg):
        for j in range(wg):
            y2, x2 = min(i * hs + hc, h_img), min(j * ws + wc, w_img)
            wins.append((max(y2 - hc, 0), max(x2 - wc, 0), y2, x2))
    return wins


def slide_inference(encode_decode, img, num_classes, crop, stride):
    N, _, H, W = img.shape
    preds = img.new_zeros((N, num_classes, H, W))
    cnt = img.new_zeros((N, 1, H, W))
    for (y1, x1, y2, x2) in slide_windows(H, W, crop, stride):
        preds[:, :, y1:y2, x1:x2] += encode_decode(img[:, :, y1:y2, x1:x2])
        cnt[:, :, y1:y2, x1:x2] += 1
    return preds / cnt


def simple_test(seg_logit):
    """softmax(dim=1) -> argmax(dim=1) (encoder_decoder.py:253,272)."""
    return torch.softmax(seg_logit, dim=1).argmax(dim=1)


# --------------------------------------------------------------------------
# integer area histograms / confusion matrix (metrics.py:26-87; tests/test_metrics.py:9-26)
# --------------------------------------------------------------------------
def confusion_matrix(pred: np.ndarray, label: np.ndarray, num_classes: int, ignore_index: int = 255) -> np.ndarray:
    """CM[label, pred] as int64, after dropping label==ignore_index and any value outside
    [0,C-1] (histc drops out-of-range values, metrics.py:79-84)."""
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    label = np.asarray(label).reshape(-1).astype(np.int64)
    keep = (label != ignore_index) & (label >= 0) & (label < num_classes) & (pred >= 0) & (pred < num_classes)
    return np.bincount(num_classes * label[keep] + pred[keep], minlength=num_classes ** 2) \
        .reshape(num_classes, num_classes).astype(np.int64)


def intersect_and_union(pred: np.ndarray, label: np.ndarray, num_classes: int, ignore_index: int = 255):
    """The four area vectors of metrics.py:75-87 as int64.  NB the reference histograms pred and
    label independently (a pixel whose pred is out of range still counts in area_label)."""
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    label = np.asarray(label).reshape(-1).astype(np.int64)
    m = label != ignore_index
    pred, label = pred[m], label[m]
    inr = lambda a: a[(a >= 0) & (a < num_classes)]
    a_i = np.bincount(inr(pred[pred == label]), minlength=num_classes)
    a_p = np.bincount(inr(pred), minlength=num_classes)
    a_l = np.bincount(inr(label), minlength=num_classes)
    return a_i.astype(np.int64), (a_p + a_l - a_i).astype(np.int64), a_p.astype(np.int64), a_l.astype(np.int64)


def metrics_from_confusion(cm: np.ndarray, beta: float = 1.0) -> Dict[str, np.ndarray]:
    """Standard (un-tampered) definitions pinned by tests/test_metrics.py:29-85."""
    cm = cm.astype(np.float64)
    tp, row, col = np.diag(cm), cm.sum(1), cm.sum(0)
    with np.errstate(divide="ignore", invalid="ignore"):
        prec, rec = tp / col, tp / row
        out = {"aAcc": tp.sum() / cm.sum(), "Acc": rec, "IoU": tp / (row + col - tp),
               "Dice": 2 * tp / (row + col), "Precision": prec, "Recall": rec,
               "Fscore": (1 + beta ** 2) * prec * rec / (beta ** 2 * prec + rec)}
    return out


# --------------------------------------------------------------------------
# family B: mmseg UNet (backbones/unet.py) + FCNHead (decode_heads/fcn_head.py)
# --------------------------------------------------------------------------
def conv_module(sd: SD, p: str, x, train, ns, k: int):
    """mmcv ConvModule: conv (no bias when a norm follows) -> BN -> ReLU (unet.py:66-75)."""
    x = F.conv2d(x, sd[p + ".conv.weight"], sd.get(p + ".conv.bias"), padding=k // 2)
    if p + ".bn.weight" in sd:
        x = batch_norm(sd, p + ".bn", x, train, ns)
    return torch.relu(x)


def basic_conv_block(sd: SD, p: str, x, train, ns):
    i = 0
    while f"{p}.convs.{i}.conv.weight" in sd:
        x = conv_module(sd, f"{p}.convs.{i}", x, train, ns, 3)
        i += 1
    return x


def unet_b_forward(sd: SD, x, train: bool = True, new_stats: Optional[dict] = None) -> List[torch.Tensor]:
    """UNet.forward (unet.py:404-415) for strides all 1 / MaxPool downsampling / InterpConv or DeconvModule upsampling:
    encoder stage i = [MaxPool2d(2)] + BasicConvBlock; decoder i = UpConvBlock(skip=enc_i, x): InterpConv (bilinear x2,
    align_corners=False, then 1x1 ConvModule) -> cat[skip, x] -> BasicConvBlock (up_conv_block.py:95-102)."""
    enc = []
    i = 0
    while any(k.startswith(f"encoder.{i}.") for k in sd):
        if i > 0:
            x = F.max_pool2d(x, 2)
        blk = f"encoder.{i}.{1 if i > 0 else 0}"
        x = basic_conv_block(sd, blk, x, train, new_stats)
        enc.append(x)
        i += 1
    outs = [x]
    for j in reversed(range(len(enc) - 1)):
        dk = f"decoder.{j}.upsample.deconv_upsamping"
        if dk + ".0.weight" in sd:   # DeconvModule (unet.py:89-147): ConvTranspose2d(4, 2, 1) -> BN -> ReLU
            up = F.conv_transpose2d(x, sd[dk + ".0.weight"], sd[dk + ".0.bias"], stride=2, padding=1)
            up = torch.relu(batch_norm(sd, dk + ".1", up, train, new_stats))
        else:
            up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
            up = conv_module(sd, f"decoder.{j}.upsample.interp_upsample.1", up, train, new_stats, 1)
        x = basic_conv_block(sd, f"decoder.{j}.conv_block", torch.cat([enc[j], up], dim=1), train, new_stats)
        outs.append(x)
    return outs


def fcn_head_forward(sd: SD, feats: List[torch.Tensor], in_index: int, train: bool = True, new_stats: Optional[dict] = None):
    """FCNHead.forward (fcn_head.py:67-88) with num_convs ConvModules, optional concat_input, then cls_seg (no dropout)."""
    x = feats[in_index]
    f = x
    i = 0
    while f"convs.{i}.conv.weight" in sd:
        f = conv_module(sd, f"convs.{i}", f, train, new_stats, sd[f"convs.{i}.conv.weight"].shape[-1])
        i += 1
    if "conv_cat.conv.weight" in sd:
        f = conv_module(sd, "conv_cat", torch.cat([x, f], dim=1), train, new_stats, sd["conv_cat.conv.weight"].shape[-1])
    return F.conv2d(f, sd["conv_seg.weight"], sd["conv_seg.bias"])


# --------------------------------------------------------------------------
# config 5: UnetPlusPlus head = smp.UnetPlusPlus(encoder_name="vgg16", classes=64) + cls_seg (unetpp_head.py:11-22)
# PARITY UNPINNED: segmentation_models_pytorch 0.2.0 is not vendored under /root/reference and not installed; this restates its
# published design (torchvision VGG16 features split at the max-pools; UNet++ nested decoder with nearest x2 up-sampling, dense
# skip concatenation [x, dense.., feature], two conv3x3(no bias)-BN-ReLU per block, decoder channels (256,128,64,32,16);
# 3x3 segmentation head).  No reference test covers it; the only anchors are the call site and the channel bookkeeping.
# --------------------------------------------------------------------------
_VGG16_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M")


def unetpp_forward(sd: SD, img, train: bool = True, new_stats: Optional[dict] = None):
    x, feats, idx = img, [], 0
    for v in _VGG16_CFG:
        if v == "M":
            feats.append(x)
            x = F.max_pool2d(x, 2)
            idx += 1
        else:
            x = torch.relu(F.conv2d(x, sd[f"model.encoder.features.{idx}.weight"], sd[f"model.encoder.features.{idx}.bias"], padding=1))
            idx += 2
    feats.append(x)
    f = feats[1:][::-1]

    def block(name, x, skips):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skips:
            x = torch.cat([x] + skips, dim=1)
        for c in ("conv1", "conv2"):
            p = f"model.decoder.blocks.{name}.{c}"
            x = torch.relu(batch_norm(sd, p + ".1", F.conv2d(x, sd[p + ".0.weight"], None, padding=1), train, new_stats))
        return x
    depth, dense = 4, {}
    for l in range(depth):
        for d in range(depth - l):
            if l == 0:
                dense[f"x_{d}_{d}"] = block(f"x_{d}_{d}", f[d], [f[d + 1]])
            else:
                li = d + l
                cat = [dense[f"x_{i}_{li}"] for i in range(d + 1, li + 1)] + [f[li + 1]]
                dense[f"x_{d}_{li}"] = block(f"x_{d}_{li}", dense[f"x_{d}_{li - 1}"], cat)
    y = block(f"x_0_{depth}", dense[f"x_0_{depth - 1}"], [])
    y = F.conv2d(y, sd["model.segmentation_head.0.weight"], sd["model.segmentation_head.0.bias"], padding=1)
    return F.conv2d(y, sd["conv_seg.weight"], sd["conv_seg.bias"])
