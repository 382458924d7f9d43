"""Pins the oracle restatement (oracle/stc_oracle.py) — CPU only.
  (a) against the reference's OWN modules imported from /root/reference (skipped where that tree is absent, e.g. the GPU box)
  (b) against the committed golden fixtures generated from those modules (oracle/make_golden.py)
  (c) against the reference's known-answer tests for CE and the confusion matrix."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_shim, stc_oracle as O
from tests.util import rel_l2

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))
LOSS_CFG = ref_shim.LOSS_CFG


def build_ours(kind, num_classes, seed):
    """Our module classes are parameter containers with the reference's construction order -> identical default init."""
    import stc_unet_b200 as S
    torch.manual_seed(seed)
    if kind == "stc":
        bb = S.build_backbone(dict(type="UnetBackbone", in_channels=3, context_layer="kernelselect", transformer_block=True,
                                   channel_list=[64, 128, 256, 512]))
        hd = S.build_head(dict(type="UnetHead", se=True, num_classes=num_classes, channels=64, threshold=0.2,
                               norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG, dropout_ratio=0.0))
    else:
        bb = S.build_backbone(dict(type="UnetBackbone", in_channels=3, channel_list=[64, 128, 256, 512]))
        hd = S.build_head(dict(type="UnetHead", num_classes=num_classes, channels=64, threshold=0.2,
                               norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG, dropout_ratio=0.0))
    bb.init_weights(); hd.init_weights()
    return bb, hd


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def as_leaf(sd):
    return {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}


@pytest.mark.skipif(not GOLDEN, reason="no golden fixtures")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_golden(path):
    fix = torch.load(path, weights_only=False)
    bb, hd = build_ours(fix["kind"], fix["num_classes"], fix["seed"])
    merged = {**{"b." + k: v for k, v in bb.state_dict().items()}, **{"h." + k: v for k, v in hd.state_dict().items()}}
    if fix["torch_version"] == torch.__version__:
        assert sd_hash(merged) == fix["init_hash"], "seeded default init differs from the reference's"
    # state_dict layout == the reference's
    keys = {**{"backbone." + k: tuple(v.shape) for k, v in bb.state_dict().items()},
            **{"decode_head." + k: tuple(v.shape) for k, v in hd.state_dict().items()}}
    assert keys == fix["state_keys"]
    g = torch.Generator().manual_seed(fix["seed"] + 100)
    img = torch.rand(fix["batch"], 3, fix["size"], fix["size"], generator=g)
    gt = torch.randint(0, fix["num_classes"], (fix["batch"], 1, fix["size"], fix["size"]), generator=g)
    gt[:, :, :2] = 255
    bsd, hsd = as_leaf(bb.state_dict()), as_leaf(hd.state_dict())
    nb, nh = {}, {}
    out = O.forward_train(bsd, hsd, img, gt, True, nb, nh)
    (out["loss_bce"] + out["loss_dice"]).backward()
    assert rel_l2(out["logits"], fix["logits"]) < 2e-5
    for k in ("loss_bce", "loss_dice", "acc_seg"):
        assert abs(float(out[k]) - fix["losses"][k]) < 1e-5 * max(1.0, abs(fix["losses"][k]))
    for k, gref in fix["grads"].items():
        sd, name = (bsd, k[len("backbone."):]) if k.startswith("backbone.") else (hsd, k[len("decode_head."):])
        assert rel_l2(sd[name].grad, gref) < 2e-3, k
    worst = 0.0
    for k, n in fix["grad_norms"].items():
        sd, name = (bsd, k[len("backbone."):]) if k.startswith("backbone.") else (hsd, k[len("decode_head."):])
        if n > 1e-4:
            worst = max(worst, abs(float(sd[name].grad.norm()) - n) / n)
    assert worst < 5e-3
    for k, v in fix["running"].items():
        assert rel_l2(nb[k[len("backbone."):]], v) < 1e-5, k
    pred = out["logits"].argmax(1).numpy()
    label = gt.squeeze(1).numpy()
    for i in range(fix["batch"]):
        got = O.intersect_and_union(pred[i], label[i], fix["num_classes"], 255)
        assert all(np.array_equal(got[j], fix["areas"][i, j].numpy()) for j in range(4))


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("stc", [False, True])
def test_oracle_matches_reference_modules(stc):
    bb, hd = ref_shim.build_reference_model(stc, 3)
    torch.manual_seed(1)
    x = torch.rand(2, 3, 48, 48)
    y = torch.randint(0, 3, (2, 1, 48, 48)); y[:, :, :5] = 255
    bsd, hsd = as_leaf(bb.state_dict()), as_leaf(hd.state_dict())
    out = O.forward_train(bsd, hsd, x, y, True, {}, {})
    (out["loss_bce"] + out["loss_dice"]).backward()
    l = hd.forward_train(bb(x), None, y, None)
    (l["loss_bce"] + l["loss_dice"]).backward()
    for k in ("loss_bce", "loss_dice", "acc_seg"):
        assert abs(float(l[k]) - float(out[k])) < 1e-5 * max(1.0, abs(float(l[k])))
    for mod, sd in ((bb, bsd), (hd, hsd)):
        for name, p in mod.named_parameters():
            if p.grad.norm() > 1e-5 and not name.endswith(("conv.0.bias", "conv.3.bias", "ca.conv1.bias")) and ".convs." not in name:
                assert rel_l2(sd[name].grad, p.grad) < 2e-2, name   # fp32 ReLU/max-pool decision flips
    # eval-mode forward and inference post-processing
    bb.eval(); hd.eval()
    with torch.no_grad():
        ref = hd.forward_test(bb(x), None, None)
        got = O.head_forward(hd.state_dict(), O.backbone_forward(bb.state_dict(), x, False), False)
    assert rel_l2(got, ref) < 1e-5
    assert torch.equal(O.simple_test(got), F.softmax(ref, dim=1).argmax(dim=1))


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
def test_area_histograms_match_reference_metrics():
    ns = ref_shim.load_reference()
    rng = np.random.RandomState(0)
    for C in (2, 3, 19):
        pred = rng.randint(0, C, size=(30, 30)); label = rng.randint(0, C, size=(30, 30)).astype(np.uint8)
        label[2, 5:10] = 255
        ref = [t.numpy().astype(np.int64) for t in ns.intersect_and_union(pred, label, C, 255)]
        got = O.intersect_and_union(pred, label, C, 255)
        assert all(np.array_equal(a, b) for a, b in zip(ref, got))
        cm = O.confusion_matrix(pred, label, C, 255)
        assert np.array_equal(np.diag(cm), got[0]) and np.array_equal(cm.sum(0), got[2]) and np.array_equal(cm.sum(1), got[3])


def test_confusion_matrix_definition():
    """tests/test_metrics.py:9-26 (np.bincount(n*label+pred)) and :29-85 (IoU / Dice / Fscore from it)."""
    rng = np.random.RandomState(1)
    C = 19
    pred = rng.randint(0, C, size=(10, 30, 30)); label = rng.randint(0, C, size=(10, 30, 30)); label[:, 2, 5:10] = 255
    mask = label != 255
    want = np.bincount(C * label[mask] + pred[mask], minlength=C * C).reshape(C, C)
    cm = O.confusion_matrix(pred, label, C, 255)
    assert np.array_equal(cm, want)
    m = O.metrics_from_confusion(cm)
    tot = want.astype(np.float64)
    assert np.allclose(m["IoU"], np.diag(tot) / (tot.sum(1) + tot.sum(0) - np.diag(tot)))
    assert np.allclose(m["Dice"], 2 * np.diag(tot) / (tot.sum(1) + tot.sum(0)))
    assert np.isclose(m["aAcc"], np.diag(tot).sum() / tot.sum())


def test_ce_known_answers():
    """tests/test_models/test_losses/test_ce_loss.py:25-39 and :43-86."""
    assert abs(float(O.cross_entropy_loss(torch.tensor([[100.0, -100.0]]).view(1, 2, 1, 1), torch.tensor([1]).view(1, 1, 1))) - 200.0) < 1e-4
    logits = torch.full((2, 21, 8, 8), 0.5)
    label = torch.ones(2, 8, 8).long(); label[:, 0, 0] = 255
    want = F.cross_entropy(logits, label, reduction="none", ignore_index=255).sum() / label.numel()
    assert abs(float(O.cross_entropy_loss(logits, label)) - float(want)) < 1e-6


def test_slide_windows_match_reference_grid():
    """encoder_decoder.py:164-179, incl. the tests' (3,3)/(2,2) on 8x16 fixture and the 256/170 default on 512x512."""
    import stc_unet_b200 as S
    for (H, W, crop, stride) in ((8, 16, (3, 3), (2, 2)), (512, 512, (256, 256), (170, 170)), (100, 90, (64, 64), (40, 40)), (30, 30, (64, 64), (40, 40))):
        hc, wc = crop; hs, ws = stride
        hg = max(H - hc + hs - 1, 0) // hs + 1; wg = max(W - wc + ws - 1, 0) // ws + 1
        want = []
        for i in range(hg):
            for j in range(wg):
                y1, x1 = i * hs, j * ws
                y2, x2 = min(y1 + hc, H), min(x1 + wc, W)
                want.append((max(y2 - hc, 0), max(x2 - wc, 0), y2, x2))
        assert O.slide_windows(H, W, crop, stride) == want == S.slide_windows(H, W, crop, stride)
    assert len(O.slide_windows(512, 512, (256, 256), (170, 170))) == 9


def build_ours_b(num_classes, base=16, stages=4, dtype="bf16", seed=0, upsample="InterpConv"):
    import stc_unet_b200 as S
    torch.manual_seed(seed)
    norm = dict(type="BN", requires_grad=True)
    bb = S.build_backbone(dict(type="UNet", in_channels=3, base_channels=base, num_stages=stages, strides=(1,) * stages,
                               enc_num_convs=(2,) * stages, dec_num_convs=(2,) * (stages - 1), downsamples=(True,) * (stages - 1),
                               enc_dilations=(1,) * stages, dec_dilations=(1,) * (stages - 1), with_cp=False, conv_cfg=None, norm_cfg=norm,
                               act_cfg=dict(type="ReLU"), upsample_cfg=dict(type=upsample), norm_eval=False, compute_dtype=dtype))
    hd = S.build_head(dict(type="FCNHead", in_channels=base, in_index=stages - 1, channels=base, num_convs=1, concat_input=False,
                           dropout_ratio=0.0, num_classes=num_classes, norm_cfg=norm, align_corners=False,
                           loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0)))
    return bb, hd


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
def test_family_b_oracle_and_state_dict_match_reference():
    """mmseg UNet-S5-D16 + FCNHead (configs/_base_/models/fcn_unet_s5-d16.py): our containers have the reference's state_dict,
    the oracle restatement reproduces the reference modules (ConvModule itself is a restatement: mmcv is not installed),
    and the output shape table of tests/test_models/test_backbones/test_unet.py holds."""
    rb, rh = ref_shim.build_reference_model_b(3, base_channels=16, num_stages=4)
    bb, hd = build_ours_b(3)
    for ours, ref in ((bb, rb), (hd, rh)):
        so, sr = ours.state_dict(), ref.state_dict()
        assert {k: tuple(v.shape) for k, v in so.items()} == {k: tuple(v.shape) for k, v in sr.items()}
        ours.load_state_dict(sr, strict=True)
    x = torch.rand(2, 3, 32, 32)
    y = torch.randint(0, 3, (2, 1, 32, 32))
    feats = rb(x)
    assert [tuple(f.shape) for f in feats] == [(2, 128, 4, 4), (2, 64, 8, 8), (2, 32, 16, 16), (2, 16, 32, 32)]
    l = rh.forward_train(feats, None, y, None)
    l["loss_ce"].backward()
    bsd, hsd = as_leaf(rb.state_dict()), as_leaf(rh.state_dict())
    logits = O.fcn_head_forward(hsd, O.unet_b_forward(bsd, x, True, {}), 3, True, {})
    out = O.losses(logits, y)
    out["loss_bce"].backward()
    assert abs(float(out["loss_bce"]) - float(l["loss_ce"])) < 1e-5
    for mod, sd in ((rb, bsd), (rh, hsd)):
        for name, p in mod.named_parameters():
            if p.grad.norm() > 1e-6:
                assert rel_l2(sd[name].grad, p.grad) < 2e-2, name
    # full-size S5-D16 shape table (test_unet.py: strides all 1 on 128x128 -> 1024@8 ... 64@128)
    b5, _ = ref_shim.build_reference_model_b(2, base_channels=4, num_stages=5)
    with torch.no_grad():
        assert [tuple(f.shape) for f in b5(torch.rand(1, 3, 64, 64))] == [(1, 64, 4, 4), (1, 32, 8, 8), (1, 16, 16, 16), (1, 8, 32, 32), (1, 4, 64, 64)]
    with pytest.raises(AssertionError):
        b5(torch.rand(1, 3, 65, 65))
    import stc_unet_b200 as S
    with pytest.raises(AssertionError):   # same argument validation as unet.py:324-355
        S.build_backbone(dict(type="UNet", num_stages=5, strides=(1, 1, 1, 1)))


def test_unetpp_structure_and_oracle_shapes():
    """Config 5 host structure (no GPU): smp key layout and channel bookkeeping, oracle restatement runs on the same state_dict."""
    import stc_unet_b200 as S
    hd = S.build_head(dict(type="UnetPlusPlus", num_classes=2, norm_cfg=dict(type="BN"), dropout_ratio=0.0,
                           loss_decode=[dict(type="CrossEntropyLoss", loss_name="loss_bce"), dict(type="DiceLoss", loss_name="loss_dice")]))
    sd = hd.state_dict()
    blocks = hd.model.decoder.blocks
    assert sorted(blocks) == sorted([f"x_{d}_{l}" for l in range(4) for d in range(l + 1)] + ["x_0_4"])
    want = {"x_0_0": (256, 1024), "x_1_1": (512, 1024), "x_0_1": (128, 1280), "x_2_2": (256, 768), "x_0_2": (64, 896),
            "x_3_3": (128, 384), "x_0_3": (32, 576), "x_0_4": (16, 32)}
    for k, (co, ci) in want.items():
        assert blocks[k].conv1[0].weight.shape == (co, ci, 3, 3), k
    assert sd["model.encoder.features.28.weight"].shape == (512, 512, 3, 3) and sd["model.segmentation_head.0.weight"].shape == (64, 16, 3, 3)
    img = torch.rand(1, 3, 32, 32)
    with torch.no_grad():
        assert O.unetpp_forward(sd, img, False).shape == (1, 2, 32, 32)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
def test_family_b_deconv_upsampler_matches_reference():
    """upsample_cfg=dict(type='DeconvModule') (unet.py:89-147): same state_dict keys as the reference module, the oracle restatement
    reproduces the reference forward/backward, and the 3x3-sub-pixel weight mapping used by the CUDA path equals conv_transpose2d."""
    rb, _ = ref_shim.build_reference_model_b(3, base_channels=16, num_stages=3, upsample="DeconvModule")
    bb, _ = build_ours_b(3, stages=3, upsample="DeconvModule")
    so, sr = bb.state_dict(), rb.state_dict()
    assert {k: tuple(v.shape) for k, v in so.items()} == {k: tuple(v.shape) for k, v in sr.items()}
    assert "decoder.0.upsample.deconv_upsamping.0.bias" in so and so["decoder.1.upsample.deconv_upsamping.0.weight"].shape == (64, 32, 4, 4)
    x = torch.rand(2, 3, 16, 16)
    feats = rb(x)
    feats[-1].square().mean().backward()
    bsd = as_leaf(rb.state_dict())
    mine = O.unet_b_forward(bsd, x, True, {})
    mine[-1].square().mean().backward()
    for a, b in zip(mine, feats):
        assert rel_l2(a, b) < 1e-5
    for name, p in rb.named_parameters():
        if p.grad.norm() > 1e-6:
            assert rel_l2(bsd[name].grad, p.grad) < 2e-2, name
    from stc_unet_b200.ops import deconv4x2_weight_as_conv3
    xx, w, b = torch.randn(2, 5, 6, 7, dtype=torch.float64), torch.randn(5, 3, 4, 4, dtype=torch.float64), torch.randn(3, dtype=torch.float64)
    z = F.conv2d(xx, deconv4x2_weight_as_conv3(w), b.repeat(4), padding=1)
    y = z.view(2, 2, 2, 3, 6, 7).permute(0, 3, 4, 1, 5, 2).reshape(2, 3, 12, 14)
    assert (y - F.conv_transpose2d(xx, w, b, stride=2, padding=1)).abs().max() < 1e-12
    import stc_unet_b200 as S
    with pytest.raises(AssertionError):
        S.modules_b.DeconvModule(8, 8, kernel_size=3)


def test_unetpp_encoder_is_pinned_to_torchvision_vgg16():
    """Config 5 (unetpp_head.py:16: smp.UnetPlusPlus(encoder_name="vgg16")).  smp 0.2.0 itself is absent, but its vgg16 encoder IS
    torchvision's `vgg16().features` (smp's VGGEncoder subclasses torchvision.models.VGG, drops the classifier, and cuts the stage
    outputs at every MaxPool2d).  This pins that half: our container has torchvision's keys / shapes / default init, and the oracle's
    encoder loop reproduces torchvision's layers stage by stage.  The nested decoder remains UNPINNED (no smp source anywhere)."""
    tv = pytest.importorskip("torchvision")
    from stc_unet_b200.modules_pp import _VGGEncoder
    torch.manual_seed(3)
    ref = tv.models.vgg16(weights=None).features
    torch.manual_seed(3)
    ours = _VGGEncoder()
    rs, os_ = ref.state_dict(), ours.features.state_dict()
    assert list(rs) == list(os_) and all(tuple(rs[k].shape) == tuple(os_[k].shape) for k in rs)
    assert [type(m).__name__ for m in ref] == [type(m).__name__ for m in ours.features]
    # oracle encoder == torchvision layers, cut at each pool (stage outputs BEFORE the pool, plus the final pooled map)
    x = torch.rand(2, 3, 64, 64)
    feats, h = [], x
    with torch.no_grad():
        for m in ref:
            if isinstance(m, torch.nn.MaxPool2d):
                feats.append(h)
            h = m(h)
        feats.append(h)
    assert [f.shape[1] for f in feats] == [64, 128, 256, 512, 512, 512] and [f.shape[-1] for f in feats] == [64, 32, 16, 8, 4, 2]
    # run the oracle with hooks: unetpp_forward's encoder part recomputed with the same weights
    sd = {f"model.encoder.features.{k}": v for k, v in rs.items()}
    got, h, idx = [], x, 0
    with torch.no_grad():
        for v in O._VGG16_CFG:
            if v == "M":
                got.append(h); h = F.max_pool2d(h, 2); idx += 1
            else:
                h = torch.relu(F.conv2d(h, sd[f"model.encoder.features.{idx}.weight"], sd[f"model.encoder.features.{idx}.bias"], padding=1)); idx += 2
        got.append(h)
    for a, b in zip(got, feats):
        assert torch.equal(a, b)
