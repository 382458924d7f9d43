"""Whole-path parity on the GPU: B200 modules vs the oracle restatement (oracle/stc_oracle.py, itself pinned
to the reference's own modules by tests/test_oracle.py) with the same state_dict and the same seeded inputs.

Ground truth is the oracle evaluated in fp64.  north-star tolerances: fp32 logits/gradients <= 1e-4 relative
L2, bf16 <= 2e-2, argmax agreement >= 99.9 %, confusion matrices bit-exact.

Gradients need care (DESIGN.md "Parity protocol"): with random-init weights every dW is a random-walk sum, so
ONE ReLU/max-pool decision that flips under a 1e-6 forward perturbation moves that channel's gradient by
~1/sqrt(#pixels) — PyTorch's own fp32 path differs from the fp64 oracle by 1e-2 on such inputs.  Hence:
  * "flip-free" configuration (BN gamma=0.25, beta=4: every pre-activation positive, ReLU acts linearly):
    every gradient must meet the north-star bound (or PyTorch-fp32's own error, when that is larger);
  * default configuration: our error vs fp64 must not exceed PyTorch-fp32's (resp. autocast-bf16's) own
    error vs fp64 by more than a small factor — i.e. we match the reference as well as it matches itself."""
import statistics

import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu

LOSS_CFG = [dict(type="CrossEntropyLoss", use_sigmoid=False, loss_name="loss_bce", loss_weight=1.0),
            dict(type="DiceLoss", loss_name="loss_dice", loss_weight=1.0)]


def build(stc: bool, num_classes: int, dtype: str, seed=0, posbn=False):
    import stc_unet_b200 as S
    torch.manual_seed(seed)
    if stc:
        bb = S.build_backbone(dict(type="UnetBackbone", in_channels=3, context_layer="kernelselect", transformer_block=True,
                                   channel_list=[64, 128, 256, 512], compute_dtype=dtype))
        hd = S.build_head(dict(type="UnetHead", se=True, num_classes=num_classes, channels=64, threshold=0.2,
                               norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG, dropout_ratio=0.0))
    else:
        bb = S.build_backbone(dict(type="UnetBackbone", in_channels=3, channel_list=[64, 128, 256, 512], compute_dtype=dtype))
        hd = S.build_head(dict(type="UnetHead", num_classes=num_classes, channels=64, threshold=0.2,
                               norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG, dropout_ratio=0.0))
    bb.init_weights(); hd.init_weights()
    g = torch.Generator().manual_seed(seed + 1)
    for m in list(bb.modules()) + list(hd.modules()):
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            if posbn:
                m.weight.data.fill_(0.25); m.bias.data.fill_(4.0)
            else:  # non-trivial BN affine parameters so their gradients are exercised
                m.weight.data = torch.rand(m.weight.shape, generator=g) + 0.5
                m.bias.data = torch.rand(m.bias.shape, generator=g) - 0.5
    return bb.cuda(), hd.cuda()


def is_bn_cancelled_bias(name):
    """Conv biases feeding a train-mode BN: true gradient is 0 (SURVEY §7 traps) — excluded from relative checks."""
    return name.endswith((".conv.0.bias", ".conv.3.bias", "ca.conv1.bias")) or (".convs." in name and name.endswith(".0.bias"))


def oracle(bb, hd, img, gt, dt, autocast=False):
    from oracle import stc_oracle as O
    conv = lambda v: v.detach().clone().to(dt) if v.is_floating_point() else v.detach().clone()
    bsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in bb.state_dict().items()}
    hsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in hd.state_dict().items()}
    nb, nh = {}, {}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits = O.head_forward(hsd, O.backbone_forward(bsd, img.to(dt), True, nb), True, nh)
    out = O.losses(logits.float() if autocast else logits, gt)
    (out["loss_bce"] + out["loss_dice"]).backward()
    grads = {("b", k): v.grad for k, v in bsd.items() if v.requires_grad}
    grads.update({("h", k): v.grad for k, v in hsd.items() if v.requires_grad})
    return dict(logits=logits.detach(), grads=grads, losses=out, new_b=nb, new_h=nh)


def ours(stc, C, dtype, img, gt, posbn):
    bb, hd = build(stc, C, dtype, posbn=posbn)
    losses = hd.forward_train(bb(img), None, gt, None)
    (losses["loss_bce"] + losses["loss_dice"]).backward()
    grads = {("b", k): p.grad for k, p in bb.named_parameters()}
    grads.update({("h", k): p.grad for k, p in hd.named_parameters()})
    assert all(g is not None for g in grads.values())
    b2, h2 = build(stc, C, dtype, posbn=posbn)
    with torch.no_grad():
        logits = h2(b2(img))
    return dict(logits=logits, grads=grads, losses=losses, bb=bb, hd=hd)


def inputs(N, C, H, W, ignore_border=True):
    g = torch.Generator().manual_seed(123)
    img = torch.rand(N, 3, H, W, generator=g).cuda()
    gt = torch.randint(0, C, (N, 1, H, W), generator=g)
    if ignore_border:
        gt[:, :, :3] = 255
    return img, gt.cuda()


def grad_errors(res, ref64):
    return {k: rel_l2(res["grads"][k], ref64["grads"][k]) for k in ref64["grads"] if not is_bn_cancelled_bias(k[1])}


def check_bn_cancelled(res, ref64, abs_tol):
    for k, g in ref64["grads"].items():
        if is_bn_cancelled_bias(k[1]):
            scale = max(1.0, float(g.abs().max()))
            assert float((res["grads"][k].double() - g).abs().max()) < abs_tol * scale, k


@pytest.mark.parametrize("stc", [False, True])
@pytest.mark.parametrize("C", [2, 3])
@pytest.mark.parametrize("posbn", [False, True])
def test_fp32_parity(stc, C, posbn):
    img, gt = inputs(2, C, 64, 64)
    bb, hd = build(stc, C, "fp32", posbn=posbn)
    ref64 = oracle(bb, hd, img, gt, torch.float64)
    ref32 = oracle(bb, hd, img, gt, torch.float32)
    got = ours(stc, C, "fp32", img, gt, posbn)
    # forward: strict north-star bound
    assert rel_l2(got["logits"], ref64["logits"]) <= 1e-4
    assert float((got["logits"].argmax(1) == ref64["logits"].argmax(1)).float().mean()) >= 0.999
    for k in ("loss_bce", "loss_dice"):
        assert abs(float(got["losses"][k]) - float(ref64["losses"][k])) <= 1e-5 * max(1.0, abs(float(ref64["losses"][k])))
    assert abs(float(got["losses"]["acc_seg"]) - float(ref64["losses"]["acc_seg"])) < 0.05
    # running statistics after one training step
    for mod, new in ((got["bb"], ref64["new_b"]), (got["hd"], ref64["new_h"])):
        sd = mod.state_dict()
        for k, v in new.items():
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v), k
            else:
                assert rel_l2(sd[k], v) <= 1e-5, k
    # gradients
    e_ours, e_torch = grad_errors(got, ref64), grad_errors(ref32, ref64)
    check_bn_cancelled(got, ref64, 1e-3)
    if posbn:
        # north-star bound, strictly, for EVERY gradient (measured: worst 2e-5).  The attention q/k projections (near-uniform
        # softmax at random init) and CoordAtt's conv1 (descriptors ~100x their fluctuation) are ill-conditioned in fp32 — torch's
        # own fp32 path is at 1-3e-4 there — and meet the bound because the fp32 path contracts CENTRED operands (ops._center_tokens)
        for k, e in e_ours.items():
            assert e <= 1e-4, (k, e, e_torch[k])
    else:
        assert statistics.median(e_ours.values()) <= max(1e-4, 1.5 * statistics.median(e_torch.values()))
        assert max(e_ours.values()) <= max(1e-4, 3.0 * max(e_torch.values()))


@pytest.mark.parametrize("stc", [False, True])
@pytest.mark.parametrize("posbn", [False, True])
def test_bf16_parity(stc, posbn):
    img, gt = inputs(2, 3, 64, 64)
    bb, hd = build(stc, 3, "fp32", posbn=posbn)
    ref64 = oracle(bb, hd, img, gt, torch.float64)
    refac = oracle(bb, hd, img, gt, torch.float32, autocast=True)   # the reference's bf16 path
    got = ours(stc, 3, "bf16", img, gt, posbn)
    e_log, e_log_ac = rel_l2(got["logits"], ref64["logits"]), rel_l2(refac["logits"], ref64["logits"])
    assert e_log <= max(2e-2, 1.25 * e_log_ac), (e_log, e_log_ac)
    if posbn:
        assert e_log <= 2e-2
    for k in ("loss_bce", "loss_dice"):
        assert abs(float(got["losses"][k]) - float(ref64["losses"][k])) <= 2e-2 * max(1.0, abs(float(ref64["losses"][k])))
    e_ours, e_ac = grad_errors(got, ref64), grad_errors(refac, ref64)
    assert statistics.median(e_ours.values()) <= max(2e-2, 1.25 * statistics.median(e_ac.values()))
    assert max(e_ours.values()) <= max(2e-2, 2.0 * max(e_ac.values()))


def test_fp32_parity_config1_unet_512():
    """BASELINE.json configs[0]: U-Net fwd+bwd, batch 2 of 3x512x512, 3 classes, fp32 (oracle evaluated on the GPU in fp64)."""
    img, gt = inputs(2, 3, 512, 512)
    bb, hd = build(False, 3, "fp32", posbn=True)
    ref64 = oracle(bb, hd, img, gt, torch.float64)
    ref32 = oracle(bb, hd, img, gt, torch.float32)
    got = ours(False, 3, "fp32", img, gt, True)
    assert rel_l2(got["logits"], ref64["logits"]) <= 1e-4
    assert float((got["logits"].argmax(1) == ref64["logits"].argmax(1)).float().mean()) >= 0.999
    e_ours, e_torch = grad_errors(got, ref64), grad_errors(ref32, ref64)
    for k, e in e_ours.items():
        assert e <= max(1e-4, 2.0 * e_torch[k]), (k, e, e_torch[k])


def logits_case(stc, C, dtype, N, H, W):
    from oracle import stc_oracle as O
    bb, hd = build(stc, C, dtype)
    g = torch.Generator().manual_seed(7)
    img = torch.rand(N, 3, H, W, generator=g).cuda()
    bsd = {k: v.double() if v.is_floating_point() else v for k, v in bb.state_dict().items()}
    hsd = {k: v.double() if v.is_floating_point() else v for k, v in hd.state_dict().items()}
    with torch.no_grad():
        ref = O.head_forward(hsd, O.backbone_forward(bsd, img.double(), True, None), True, None)
        out = hd(bb(img))
    return out, ref


def test_odd_size_pad_path():
    """Inputs not divisible by 16 exercise Up.forward's F.pad branch (unet_head.py:52-54)."""
    out, ref = logits_case(False, 2, "fp32", 1, 72, 88)
    assert rel_l2(out, ref) <= 1e-4


def test_eval_mode_and_slide_inference():
    import numpy as np
    import stc_unet_b200 as S
    from oracle import stc_oracle as O
    bb, hd = build(True, 3, "fp32")
    seg = S.EncoderDecoder(bb, hd, test_cfg=dict(mode="slide", crop_size=(64, 64), stride=(40, 40))).cuda().eval()
    g = torch.Generator().manual_seed(11)
    img = torch.rand(2, 3, 96, 112, generator=g).cuda()
    bsd = {k: v.double() if v.is_floating_point() else v for k, v in bb.state_dict().items()}
    hsd = {k: v.double() if v.is_floating_point() else v for k, v in hd.state_dict().items()}
    enc = lambda t: O.head_forward(hsd, O.backbone_forward(bsd, t, False, None), False, None)
    with torch.no_grad():
        ref_logits = O.slide_inference(enc, img.double(), 3, (64, 64), (40, 40))
        ref_pred = O.simple_test(ref_logits)
    got_logits = seg.slide_inference(img)
    assert rel_l2(got_logits, ref_logits) <= 1e-4
    pred = seg.inference_device(img)
    assert float((pred == ref_pred).float().mean()) >= 0.999
    res = seg.simple_test(img)
    assert isinstance(res, list) and res[0].dtype == np.int64 and res[0].shape == (96, 112)
    # confusion matrix of OUR prediction: device histogram == numpy bincount, bit exact
    label = torch.randint(0, 3, (2, 96, 112), generator=g).to(torch.uint8)
    label[:, :4] = 255
    cm, areas = S.ops.confusion_hist(pred, label.cuda(), 3, 255)
    assert np.array_equal(cm.cpu().numpy(), O.confusion_matrix(pred.cpu().numpy(), label.numpy(), 3, 255))
    want = O.intersect_and_union(pred.cpu().numpy(), label.numpy(), 3, 255)
    assert all(np.array_equal(areas.cpu().numpy()[i], want[i]) for i in range(4))


def test_dropout_training_path():
    """Dropout2d(0.1) before conv_seg (decode_head.py:132-133,256-257): with an explicit mask the result equals the
    oracle evaluated with the same mask."""
    from oracle import stc_oracle as O
    bb, hd = build(False, 3, "fp32")
    hd.dropout_ratio = 0.1
    hd.dropout = torch.nn.Dropout2d(0.1)
    g = torch.Generator().manual_seed(5)
    img = torch.rand(2, 3, 32, 32, generator=g).cuda()
    torch.manual_seed(99)
    out = hd(bb(img))
    torch.manual_seed(99)
    keep = (torch.rand((2, 64), device="cuda") >= 0.1).float() / 0.9
    bb2, hd2 = build(False, 3, "fp32")
    with torch.no_grad():
        ref = O.head_forward(hd2.state_dict(), O.backbone_forward(bb2.state_dict(), img, True, None), True, None,
                             dropout_mask=keep.view(2, 64, 1, 1))
    assert rel_l2(out, ref) <= 1e-4


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_trainer_step_matches_plain_backward(dtype):
    """Trainer (flat params, gradient arena, batched weight packing + workspace arena in replay mode, fused Adam with lr=0)
    must produce the same gradients as a plain loss.backward() through the same modules."""
    import stc_unet_b200 as S
    from stc_unet_b200.train import Trainer
    img, gt = inputs(2, 3, 64, 64)
    # flip-free BN configuration: fp32 atomics make two runs differ by ~1e-7, which must not flip ReLU decisions
    bb, hd = build(True, 3, dtype, posbn=True)
    losses = hd.forward_train(bb(img), None, gt, None)
    (losses["loss_bce"] + losses["loss_dice"]).backward()
    ref = {("b", k): p.grad.clone() for k, p in bb.named_parameters()}
    ref.update({("h", k): p.grad.clone() for k, p in hd.named_parameters()})
    b2, h2 = build(True, 3, dtype, posbn=True)
    seg = S.EncoderDecoder(b2, h2).cuda().train()
    tr = Trainer(seg, lr=0.0)
    for _ in range(3):          # step 1 records, step 2 finalises + replays, step 3 is steady state
        lv = tr.step(img, gt)
    assert tr.cache.mode == "replay" and len(tr.cache.order) > 100
    got = {("b", k): p.grad for k, p in b2.named_parameters()}
    got.update({("h", k): p.grad for k, p in h2.named_parameters()})
    errs = []
    for k, g in ref.items():
        if is_bn_cancelled_bias(k[1]):
            assert float(got[k].abs().max()) == 0.0
            continue
        assert got[k].data_ptr() >= tr.arena.flat.data_ptr() and got[k].data_ptr() < tr.arena.flat.data_ptr() + tr.arena.flat.numel() * 4
        errs.append(rel_l2(got[k], g))
    if dtype == "fp32":
        assert max(errs) <= 5e-4 and statistics.median(errs) <= 1e-5
    else:
        # bf16 runs are not bit-reproducible (fp32 atomics in the KSA pooling feed bf16 roundings and ReLU decisions);
        # two runs of the SAME path differ by a few percent in the deep gradients, so only the bulk is compared
        assert statistics.median(errs) <= 0.1
    assert abs(float(lv["loss"]) - float(losses["loss_bce"] + losses["loss_dice"])) < 1e-3


@pytest.mark.parametrize("upsample", ["InterpConv", "DeconvModule"])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_family_b_unet_fcn_parity(dtype, upsample):
    """mmseg UNet + FCNHead (family B): InterpConv (bilinear align_corners=False + 1x1 ConvModule) or DeconvModule
    (ConvTranspose2d 4/2/1 + BN + ReLU) up-samplers, channel concat, bias-free ConvModules, CE-only loss — against the fp64 oracle."""
    from oracle import stc_oracle as O
    from tests.test_oracle import build_ours_b
    bb, hd = build_ours_b(3, base=64, stages=4, dtype=dtype, upsample=upsample)
    bb.init_weights(); hd.init_weights()
    bb, hd = bb.cuda(), hd.cuda()
    img, gt = inputs(2, 3, 64, 64)
    conv = lambda v: v.detach().double() if v.is_floating_point() else v.detach().clone()
    bsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in bb.state_dict().items()}
    hsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in hd.state_dict().items()}
    ref_logits = O.fcn_head_forward(hsd, O.unet_b_forward(bsd, img.double(), True, None), 3, True, None)
    ref = O.losses(ref_logits, gt)
    ref["loss_bce"].backward()
    feats = bb(img)
    assert [tuple(f.shape) for f in feats] == [(2, 512, 8, 8), (2, 256, 16, 16), (2, 128, 32, 32), (2, 64, 64, 64)]
    losses = hd.forward_train(feats, None, gt, None)
    assert set(losses) == {"loss_ce", "acc_seg"}
    losses["loss_ce"].backward()
    tol = 1e-4 if dtype == "fp32" else 3e-2
    assert abs(float(losses["loss_ce"]) - float(ref["loss_bce"])) <= tol
    errs = []
    for mod, sd in ((bb, bsd), (hd, hsd)):
        for name, p in mod.named_parameters():
            assert p.grad is not None, name
            errs.append(rel_l2(p.grad, sd[name].grad))
    assert statistics.median(errs) <= (2e-2 if dtype == "fp32" else 0.6)   # random-init decision flips, see module docstring
    with torch.no_grad():
        b2, h2 = build_ours_b(3, base=64, stages=4, dtype=dtype, upsample=upsample)
        b2.init_weights(); h2.init_weights()
        out = h2.cuda()(b2.cuda()(img))
    if dtype == "fp32":
        assert rel_l2(out, ref_logits) <= 1e-4
    else:   # the reference's bf16 path (autocast) as the yardstick
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            ac = O.fcn_head_forward(h2.state_dict(), O.unet_b_forward(b2.state_dict(), img, True, None), 3, True, None).float()
        assert rel_l2(out, ref_logits) <= max(2e-2, 1.25 * rel_l2(ac, ref_logits))


def test_slide_inference_config4_512_and_confusion_meter():
    """BASELINE.json configs[3]: test_cfg mode='slide', crop 256 / stride 170 on 512x512 slices (9 windows per slice, all
    windows of the batch in one forward), argmax, and the integer confusion matrix — vs the fp64 oracle, bit-exact histogram."""
    import numpy as np
    import stc_unet_b200 as S
    from oracle import stc_oracle as O
    from stc_unet_b200.metrics import ConfusionMeter
    bb, hd = build(True, 3, "fp32")
    seg = S.EncoderDecoder(bb, hd, test_cfg=dict(mode="slide", crop_size=(256, 256), stride=(170, 170))).cuda().eval()
    g = torch.Generator().manual_seed(21)
    img = torch.rand(2, 3, 512, 512, generator=g).cuda()
    label = torch.randint(0, 3, (2, 512, 512), generator=g).to(torch.uint8)
    label[:, :7] = 255
    assert len(S.slide_windows(512, 512, (256, 256), (170, 170))) == 9
    bsd = {k: v.double() if v.is_floating_point() else v for k, v in bb.state_dict().items()}
    hsd = {k: v.double() if v.is_floating_point() else v for k, v in hd.state_dict().items()}
    enc = lambda t: O.head_forward(hsd, O.backbone_forward(bsd, t, False, None), False, None)
    with torch.no_grad():
        ref_pred = O.simple_test(O.slide_inference(enc, img.double(), 3, (256, 256), (170, 170)))
    pred = seg.inference_device(img)
    assert float((pred == ref_pred).float().mean()) >= 0.999
    meter = ConfusionMeter(3, 255)
    meter.update(pred[0], label[0].cuda())
    meter.update(pred[1], label[1].cuda())
    want = O.confusion_matrix(pred.cpu().numpy(), label.numpy(), 3, 255)
    assert np.array_equal(meter.cm.cpu().numpy(), want)
    per_image = [O.intersect_and_union(pred[i].cpu().numpy(), label[i].numpy(), 3, 255) for i in range(2)]
    for j, a in enumerate(meter.pre_eval_tuple()):
        assert np.array_equal(a.numpy(), per_image[0][j] + per_image[1][j])
    res = meter.compute(["mIoU", "mDice"])
    ref_m = O.metrics_from_confusion(want)
    assert abs(res["mIoU"] - float(np.nanmean(ref_m["IoU"]))) < 1e-12 and abs(res["aAcc"] - float(ref_m["aAcc"])) < 1e-12


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_unetpp_config5_parity(dtype):
    """my_config/UNet++.py: EncoderDecoderFull + UnetPlusPlus (VGG16 encoder + nested dense-skip decoder).  The oracle is a
    restatement of smp 0.2.0's published design (parity UNPINNED: smp is neither vendored nor installed)."""
    import stc_unet_b200 as S
    from oracle import stc_oracle as O
    torch.manual_seed(0)
    seg = S.build_segmentor(dict(type="EncoderDecoderFull", decode_head=dict(
        type="UnetPlusPlus", num_classes=2, norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG, dropout_ratio=0.0,
        compute_dtype=dtype))).cuda().train()
    hd = seg.decode_head
    keys = list(hd.state_dict())
    assert "model.encoder.features.0.weight" in keys and "model.decoder.blocks.x_0_4.conv2.1.running_var" in keys
    assert "model.segmentation_head.0.bias" in keys and hd.model.decoder.blocks["x_0_2"].conv1[0].weight.shape == (64, 128 + 768, 3, 3)
    img, gt = inputs(2, 2, 64, 64)
    conv = lambda v: v.detach().double() if v.is_floating_point() else v.detach().clone()
    sd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in hd.state_dict().items()}
    ref_logits = O.unetpp_forward(sd, img.double(), True, None)
    ref = O.losses(ref_logits, gt)
    (ref["loss_bce"] + ref["loss_dice"]).backward()
    out = seg.train_step(dict(img=img, img_metas=None, gt_semantic_seg=gt))
    out["loss"].backward()
    tol = 1e-4 if dtype == "fp32" else 3e-2
    assert abs(float(out["loss"]) - float(ref["loss_bce"] + ref["loss_dice"])) <= tol
    errs = []
    for name, p in hd.named_parameters():
        assert p.grad is not None, name
        if name.endswith("conv1.0.weight") or name.endswith("conv2.0.weight") or "encoder" in name or "segmentation_head" in name:
            errs.append(rel_l2(p.grad, sd[name].grad))
    assert statistics.median(errs) <= (2e-2 if dtype == "fp32" else 0.6)
    seg.eval()
    with torch.no_grad():
        logits = seg.encode_decode(img)
        ref_eval = O.unetpp_forward({k: conv(v) for k, v in hd.state_dict().items()}, img.double(), False, None)
    assert logits.shape == (2, 2, 64, 64)
    assert rel_l2(logits, ref_eval) <= (1e-4 if dtype == "fp32" else 8e-2)


def test_uint8_slide_inference_matches_float_path():
    """Slide-mode inference fed with decoded uint8 HWC slices (windows cropped from the 8-bit volume, normalised on the device) gives
    the prediction of the float NCHW interface on the same pixel values."""
    import stc_unet_b200 as S
    bb, hd = build(True, 3, "bf16")
    seg = S.EncoderDecoder(bb, hd, test_cfg=dict(mode="slide", crop_size=(64, 64), stride=(42, 42))).cuda().eval()
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (2, 128, 128, 3), dtype=torch.uint8, generator=g).cuda()
    seg.backbone.img_norm_cfg = dict(mean=[0.0], std=[255.0], to_rgb=False)
    pred_u8 = seg.inference_device(u8)
    pred_f = seg.inference_device(u8.permute(0, 3, 1, 2).float() / 255.0)
    assert pred_u8.shape == (2, 128, 128) and float((pred_u8 == pred_f).float().mean()) >= 0.999


def test_uint8_input_pipeline_matches_float_path():
    """decoded uint8 HWC pixels + uint8 labels, normalised on the device, give exactly the losses and gradients of the
    reference interface (float NCHW image normalised on the host, int64 labels)."""
    import stc_unet_b200 as S
    bb, hd = build(True, 3, "bf16")
    seg = S.EncoderDecoder(bb, hd).cuda().train()
    cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
    seg.backbone.img_norm_cfg = cfg
    g = torch.Generator().manual_seed(5)
    raw = torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8)          # BGR HWC as cv2 decodes it
    lab = torch.randint(0, 3, (2, 1, 64, 64), generator=g, dtype=torch.uint8)
    lab[:, :, :2] = 255
    mean, std = torch.tensor(cfg["mean"]), torch.tensor(cfg["std"])
    ref_img = ((raw.flip(-1).float() - mean) / std).permute(0, 3, 1, 2).contiguous()     # mmcv.imnormalize + ImageToTensor
    outs = []
    for img, gt in ((ref_img.cuda(), lab.long().cuda()), (raw.cuda(), lab.cuda())):
        seg.zero_grad(set_to_none=True)
        torch.manual_seed(11)                                                            # same Dropout2d mask
        out = seg.train_step(dict(img=img, img_metas=None, gt_semantic_seg=gt))
        out["loss"].backward()
        outs.append((float(out["loss"]), seg.backbone.inc.conv.conv[0].weight.grad.clone(), seg.decode_head.conv_seg.weight.grad.clone()))
    assert abs(outs[0][0] - outs[1][0]) <= 2e-3 * abs(outs[0][0])
    assert rel_l2(outs[1][2], outs[0][2]) < 2e-2 and rel_l2(outs[1][1], outs[0][1]) < 0.2   # bf16 input rounding differs by <= 1 ulp


def test_trainer_cuda_graph_matches_eager():
    """Trainer.capture / step_graph: the captured whole-step graph (fwd + loss + bwd + device-state Adam) trains like the eager
    step on the same data sequence (dropout off; wgrad reductions use atomics, so agreement is close, not bitwise)."""
    import stc_unet_b200 as S
    from stc_unet_b200.train import Trainer
    g = torch.Generator().manual_seed(3)
    imgs = [torch.rand(2, 3, 64, 64, generator=g).cuda() for _ in range(4)]
    gts = [torch.randint(0, 3, (2, 1, 64, 64), generator=g).cuda() for _ in range(4)]
    losses = {}
    for mode in ("eager", "graph"):
        bb, hd = build(True, 3, "fp32", posbn=True)
        seg = S.EncoderDecoder(bb, hd).cuda().train()
        tr = Trainer(seg, lr=1e-3)
        seq = []
        for s in range(3):                                   # StepCache warm-up: real steps in both modes
            seq.append(float(tr.step(imgs[s % 4], gts[s % 4])["loss"]))
        if mode == "graph":
            tr.capture(imgs[3], gts[3])                      # note: capture() itself performs one more real step on these inputs
        else:
            tr.step(imgs[3], gts[3])
        for s in range(4, 10):
            lv = tr.step_graph(imgs[s % 4], gts[s % 4]) if mode == "graph" else tr.step(imgs[s % 4], gts[s % 4])
            seq.append(float(lv["loss"]))
        losses[mode] = seq
        if mode == "graph":
            assert tr.optim.t == 10 and int(tr.optim.step_count_dev) == 10
    for a, b in zip(losses["eager"], losses["graph"]):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (losses["eager"], losses["graph"])
    assert losses["graph"][-1] < losses["graph"][0]          # it does train


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_eval_bn_folding_matches_unfolded(dtype):
    """Inference folds eval-mode BN into the conv (stc_bn_fold_conv + activation in the conv epilogue); logits must agree with the
    unfolded conv -> BN-apply path (fp32: 1e-5; bf16: the two differ by one rounding of the conv output) and with the fp64 oracle."""
    import stc_unet_b200 as S
    from oracle import stc_oracle as O
    from stc_unet_b200 import ops
    bb, hd = build(True, 3, dtype)
    g = torch.Generator().manual_seed(4)
    for m in list(bb.modules()) + list(hd.modules()):      # non-trivial running statistics
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    bb.eval(); hd.eval()
    img = torch.rand(2, 3, 64, 64, generator=g).cuda()
    with torch.no_grad():
        folded = hd(bb(img))
        ops.config.fold_eval_bn = False
        try:
            plain = hd(bb(img))
        finally:
            ops.config.fold_eval_bn = True
        bsd = {k: v.double() if v.is_floating_point() else v for k, v in bb.state_dict().items()}
        hsd = {k: v.double() if v.is_floating_point() else v for k, v in hd.state_dict().items()}
        ref = O.head_forward(hsd, O.backbone_forward(bsd, img.double(), False, None), False, None)
    if dtype == "fp32":
        assert rel_l2(folded, plain) <= 1e-5 and rel_l2(folded, ref) <= 1e-4
    else:
        assert rel_l2(folded, ref) <= max(2e-2, 1.25 * rel_l2(plain, ref))
    assert float((folded.argmax(1) == ref.argmax(1)).float().mean()) >= (0.999 if dtype == "fp32" else 0.97)


def test_eval_weight_cache_is_opt_in_and_invalidated():
    """config.cache_eval_weights keeps folded / packed operands between no-grad forwards (frozen weights); it is off by default, and
    after a weight change + ops.invalidate_weight_caches() the next forward sees the new weights."""
    import stc_unet_b200 as S
    from stc_unet_b200 import ops
    assert ops.config.cache_eval_weights is False
    bb, hd = build(True, 3, "bf16")      # STC-UNet: conv+BN layers (fold cache) and Linear / attention projections (pack cache)
    bb.eval(); hd.eval()
    img = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(2)).cuda()
    ops.config.cache_eval_weights = True
    try:
        with torch.no_grad():
            a = hd(bb(img))
            b = hd(bb(img))                       # served from the caches
            assert torch.equal(a, b) and len(ops._EVAL_FOLD_CACHE) > 0 and len(ops._INFER_PACK_CACHE) > 0
            bb.inc.conv.conv[0].weight.data.mul_(1.5)     # raw .data update: invisible to version counters
            ops.invalidate_weight_caches()
            c = hd(bb(img))
            assert not torch.equal(a, c)
    finally:
        ops.config.cache_eval_weights = False
        ops.invalidate_weight_caches()


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[1] at its real spatial size: the path the headline number runs
# ------------------------------------------------------------------------------------------------
def _trainer_grads(bb, hd):
    got = {("b", k): p.grad for k, p in bb.named_parameters()}
    got.update({("h", k): p.grad for k, p in hd.named_parameters()})
    return got


@pytest.mark.parametrize("posbn", [True, False])
def test_bf16_parity_config2_512(posbn):
    """my_config/STC-UNet.py fwd+bwd in bf16 on 512x512 slices (N = 2 here; BASELINE.json configs[1] uses N = 16 of the same
    shape), driven the way bench.py drives it: Trainer.step x 3 (StepCache record -> finalize -> replay) and then the captured
    whole-step CUDA graph, against the oracle evaluated on the GPU in fp64 and - the reference's own bf16 path - under autocast.
    At W >= 128 the convolutions run on the halo tcgen05 kernels (umma_convh / umma_wgradh), the 3x3 / 5x5 / 7x7 layers that
    qualify get their BN statistics from the conv epilogue, the image conv runs through im2col: all asserted below."""
    import stc_unet_b200 as S
    from stc_unet_b200 import ops
    from stc_unet_b200.train import Trainer
    img, gt = inputs(2, 3, 512, 512)
    bb0, hd0 = build(True, 3, "fp32", posbn=posbn)
    ref64 = oracle(bb0, hd0, img, gt, torch.float64)
    refac = oracle(bb0, hd0, img, gt, torch.float32, autocast=True)
    torch.cuda.empty_cache()
    # forward logits (train-mode BN) of the bf16 path
    b1, h1 = build(True, 3, "bf16", posbn=posbn)
    with torch.no_grad():
        logits = h1(b1(img))
    e_log, e_log_ac = rel_l2(logits, ref64["logits"]), rel_l2(refac["logits"], ref64["logits"])
    assert e_log <= max(2e-2, 1.25 * e_log_ac), (e_log, e_log_ac)
    if posbn:
        assert e_log <= 2e-2, e_log
        assert float((logits.argmax(1) == ref64["logits"].argmax(1)).float().mean()) >= 0.97
    # the training step as the bench runs it
    b2, h2 = build(True, 3, "bf16", posbn=posbn)
    seg = S.EncoderDecoder(b2, h2).cuda().train()
    tr = Trainer(seg, lr=0.0)
    names, orig = [], S._lib.lib.call
    prof = ops.LaunchProfiler(time_dense=True)
    for it in range(3):
        if it == 2:                                   # steady state (replay): record which kernels serve it
            ops.set_profiler(prof)
            inner = S._lib.lib.call
            S._lib.lib.call = lambda name, *a, _i=inner: (names.append(name), _i(name, *a))[1]
        lv = tr.step(img, gt)
    ops.set_profiler(None)
    if "call" in S._lib.lib.__dict__:
        del S._lib.lib.__dict__["call"]
    assert tr.cache.mode == "replay"
    engines = {(kind, eng) for (kind, eng) in prof.summary()}
    assert ("conv_fprop", 3) in engines, engines      # stc::umma_convh_kernel (halo fprop / dgrad)
    assert ("conv_wgrad", 4) in engines, engines      # stc::umma_wgradh_kernel (halo wgrad)
    assert ("gemm", 2) in engines, engines            # stc::umma_kernel (attention / folded linears)
    assert "stc_conv_fprop_bnstats" in names and "stc_im2col" in names and "stc_pack_conv_weights_batched" in names
    ref_loss = float(ref64["losses"]["loss_bce"] + ref64["losses"]["loss_dice"])
    assert abs(float(lv["loss"]) - ref_loss) <= 2e-2 * max(1.0, abs(ref_loss))

    def check(got, what):
        e_ours, e_ac = grad_errors(dict(grads=got), ref64), grad_errors(refac, ref64)
        med, med_ac = statistics.median(e_ours.values()), statistics.median(e_ac.values())
        assert med <= max(2e-2, 1.25 * med_ac), (what, med, med_ac)
        assert max(e_ours.values()) <= max(2e-2, 2.0 * max(e_ac.values())), (what, max(e_ours.values()), max(e_ac.values()))
        # (bf16 GRADIENTS sit at ~5e-2 of the fp64 oracle for the reference's own autocast path too, flip-free or not: 40 layers of
        # 2^-9 roundings on the way back; the north-star 2e-2 is asserted on the logits above and on every fp32 gradient elsewhere)
        report[what] = dict(grad_err_median=med, grad_err_median_autocast=med_ac, grad_err_max=max(e_ours.values()),
                            grad_err_max_autocast=max(e_ac.values()))
    report = dict(posbn=posbn, logits_err=e_log, logits_err_autocast=e_log_ac)
    check(_trainer_grads(b2, h2), "eager replay")
    # the captured graph on the same data: same gradients (lr = 0, so the weights have not moved)
    tr.capture(img, gt)
    lvg = tr.step_graph(img, gt)
    torch.cuda.synchronize()
    assert abs(float(lvg["loss"]) - ref_loss) <= 2e-2 * max(1.0, abs(ref_loss))
    check({k: g.clone() for k, g in _trainer_grads(b2, h2).items()}, "cuda graph")
    # running statistics after the steps: the first step's update is the oracle's (momentum 0.1 from the initial buffers)
    b3, h3 = build(True, 3, "bf16", posbn=posbn)
    h3.forward_train(b3(img), None, gt, None)
    for mod, new in ((b3, ref64["new_b"]), (h3, ref64["new_h"])):
        sd = mod.state_dict()
        for k, v in new.items():
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v), k
            else:
                assert rel_l2(sd[k], v) <= 2e-2, (k, rel_l2(sd[k], v))
    import json, os
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):          # measured errors kept beside the run (copied to profiles/ by hand)
        with open(os.path.join(out_dir, f"parity_config2_512_{'flipfree' if posbn else 'default'}.json"), "w") as f:
            json.dump(report, f, indent=1)


def _set_posbn(*mods):
    for mod in mods:
        for m in mod.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.weight.data.fill_(0.25); m.bias.data.fill_(4.0)


@pytest.mark.parametrize("upsample", ["InterpConv", "DeconvModule"])
def test_family_b_flip_free_strict_gradients(upsample):
    """Family B (mmseg UNet + FCNHead) in the flip-free configuration (every BN gamma = 0.25, beta = 4: all pre-activations
    positive, so no ReLU decision can flip under rounding): EVERY fp32 parameter gradient within the north-star 1e-4."""
    from oracle import stc_oracle as O
    from tests.test_oracle import build_ours_b
    bb, hd = build_ours_b(3, base=64, stages=4, dtype="fp32", upsample=upsample)
    bb.init_weights(); hd.init_weights()
    _set_posbn(bb, hd)
    bb, hd = bb.cuda(), hd.cuda()
    img, gt = inputs(2, 3, 64, 64)
    conv = lambda v: v.detach().double() if v.is_floating_point() else v.detach().clone()
    bsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in bb.state_dict().items()}
    hsd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in hd.state_dict().items()}
    ref_logits = O.fcn_head_forward(hsd, O.unet_b_forward(bsd, img.double(), True, None), 3, True, None)
    O.losses(ref_logits, gt)["loss_bce"].backward()
    losses = hd.forward_train(bb(img), None, gt, None)
    losses["loss_ce"].backward()
    worst = 0.0
    for mod, sd in ((bb, bsd), (hd, hsd)):
        for name, p in mod.named_parameters():
            g = sd[name].grad
            if float(g.norm()) < 1e-12 * max(1.0, float(sd[name].detach().norm())):     # exactly-cancelled (bias into train-mode BN)
                assert float(p.grad.abs().max()) <= 1e-6, name
                continue
            e = rel_l2(p.grad, g)
            worst = max(worst, e)
            assert e <= 1e-4, (name, e)
    assert worst > 0.0


def test_unetpp_flip_free_strict_gradients():
    """UNet++ (config 5) with flip-free decoder BNs and large positive VGG biases (the encoder has no BN; bias 4 keeps every
    encoder pre-activation positive): every fp32 gradient within 1e-4 of the (restated, UNPINNED) oracle."""
    import stc_unet_b200 as S
    from oracle import stc_oracle as O
    torch.manual_seed(0)
    seg = S.build_segmentor(dict(type="EncoderDecoderFull", decode_head=dict(
        type="UnetPlusPlus", num_classes=2, norm_cfg=dict(type="BN", requires_grad=True), loss_decode=LOSS_CFG, dropout_ratio=0.0,
        compute_dtype="fp32"))).cuda().train()
    hd = seg.decode_head
    _set_posbn(hd)
    for m in hd.model.encoder.features:
        if isinstance(m, torch.nn.Conv2d):
            m.bias.data.fill_(4.0)
    img, gt = inputs(2, 2, 64, 64)
    conv = lambda v: v.detach().double() if v.is_floating_point() else v.detach().clone()
    sd = {k: conv(v).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in hd.state_dict().items()}
    ref = O.losses(O.unetpp_forward(sd, img.double(), True, None), gt)
    (ref["loss_bce"] + ref["loss_dice"]).backward()
    out = seg.train_step(dict(img=img, img_metas=None, gt_semantic_seg=gt))
    out["loss"].backward()
    assert abs(float(out["loss"]) - float(ref["loss_bce"] + ref["loss_dice"])) <= 1e-5
    for name, p in hd.named_parameters():
        e = rel_l2(p.grad, sd[name].grad)
        assert e <= 1e-4, (name, e)


def test_gradient_accumulation_without_zero_grad():
    """Two backward passes without zero_grad (mmcv's GradientCumulativeOptimizerHook) and a weight used twice in one graph must
    ACCUMULATE, although every parameter gradient normally is a view into the flat arena our kernels overwrite."""
    import stc_unet_b200 as S
    from stc_unet_b200 import ops
    bb, hd = build(False, 3, "fp32", posbn=True)
    seg = S.EncoderDecoder(bb, hd).cuda().train()
    params = [p for p in seg.parameters() if p.requires_grad]
    img1, gt1 = inputs(2, 3, 32, 32)
    g = torch.Generator().manual_seed(77)
    img2 = torch.rand(2, 3, 32, 32, generator=g).cuda()
    gt2 = torch.randint(0, 3, (2, 1, 32, 32), generator=g).cuda()

    def grads_of(img, gt):
        seg.zero_grad(set_to_none=True)
        seg.train_step(dict(img=img, img_metas=None, gt_semantic_seg=gt))["loss"].backward()
        return [p.grad.clone() for p in params]
    g1, g2 = grads_of(img1, gt1), grads_of(img2, gt2)
    arena = ops.GradArena(params)
    ops.set_grad_arena(arena)
    try:
        seg.zero_grad(set_to_none=True)
        seg.train_step(dict(img=img1, img_metas=None, gt_semantic_seg=gt1))["loss"].backward()
        assert all(p.grad.data_ptr() == arena.view(p).data_ptr() for p in params)       # first pass: straight into the arena
        seg.train_step(dict(img=img2, img_metas=None, gt_semantic_seg=gt2))["loss"].backward()
        for p, a, b in zip(params, g1, g2):
            assert rel_l2(p.grad, a + b) <= 1e-5 or float((a + b).abs().max()) < 1e-9
            assert rel_l2(arena.view(p), a + b) <= 1e-5 or float((a + b).abs().max()) < 1e-9    # ... and the sum lives in the arena
        # one Linear weight used twice in ONE graph
        w = torch.nn.Parameter(torch.randn(64, 64, device="cuda") * 0.1)
        ar2 = ops.GradArena([w])
        ops.set_grad_arena(ar2)
        x = torch.randn(1, 40, 64, device="cuda")
        y = ops.linear_tokens(ops.linear_tokens(x, w), w)
        y.sum().backward()
        wr = w.detach().clone().requires_grad_(True)
        (x @ wr.t() @ wr.t()).sum().backward()
        assert rel_l2(w.grad, wr.grad) <= 1e-5
    finally:
        ops.set_grad_arena(None)


def test_device_augmentation_feeds_trainer_and_cuda_graph():
    """augment_batch_u8 -> Trainer.capture / step_graph: the NHWC layout of the pre-normalised activations must survive the clone()
    and copy_() the trainer applies to its inputs (it travels with the tensor TYPE), and a mis-laid-out input must raise."""
    import stc_unet_b200 as S
    from stc_unet_b200 import ops
    from stc_unet_b200.train import Trainer
    g = torch.Generator().manual_seed(9)
    raw = torch.randint(0, 256, (2, 80, 96, 3), dtype=torch.uint8, generator=g).cuda()
    lab = torch.randint(0, 3, (2, 80, 96), dtype=torch.uint8, generator=g).cuda()
    geom = torch.tensor([[3, 5, 0], [10, 20, 1]], dtype=torch.int32)
    cfg = dict(mean=[0.0], std=[255.0], to_rgb=False)
    x, y = ops.augment_batch_u8(raw, lab, geom, (64, 64), torch.float32, cfg)
    assert isinstance(x, ops.NHWCImage) and isinstance(x.clone(), ops.NHWCImage) and tuple(x.shape) == (2, 64, 64, 3)
    losses = {}
    for mode in ("eager", "graph"):
        bb, hd = build(False, 3, "fp32", posbn=True)
        seg = S.EncoderDecoder(bb, hd).cuda().train()
        tr = Trainer(seg, lr=1e-3)
        seq = [float(tr.step(x, y)["loss"]) for _ in range(3)]
        if mode == "graph":
            tr.capture(x, y)
            assert isinstance(tr._static_img, ops.NHWCImage)
        else:
            tr.step(x, y)
        for _ in range(3):
            lv = tr.step_graph(x, y) if mode == "graph" else tr.step(x, y)
            seq.append(float(lv["loss"]))
        losses[mode] = seq
    for a, b in zip(losses["eager"], losses["graph"]):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), losses
    # the same activations WITHOUT the type: read as NCHW (N, C=64, ...) -> must raise, not compute garbage
    bb, hd = build(False, 3, "fp32")
    with pytest.raises(RuntimeError):
        hd.forward_train(bb(x.as_subclass(torch.Tensor)), None, y, None)
