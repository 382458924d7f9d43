"""Whole-path parity on the GPU: B200 modules vs the oracle restatement (oracle/stc_oracle.py, itself pinned
to the reference's own modules by tests/test_oracle.py) with the same state_dict and the same seeded inputs.

north-star tolerances: fp32 logits/gradients <= 1e-4 relative L2, bf16 <= 2e-2, argmax agreement >= 99.9 %."""
import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu

LOSS_CFG = [dict(type="CrossEntropyLoss", use_sigmoid=False, loss_name="loss_bce", loss_weight=1.0),
            dict(type="DiceLoss", loss_name="loss_dice", loss_weight=1.0)]


def build(stc: bool, num_classes: int, dtype: str, seed=0):
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
    # non-trivial BN affine parameters so their gradients are exercised
    g = torch.Generator().manual_seed(seed + 1)
    for m in list(bb.modules()) + list(hd.modules()):
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.weight.data = torch.rand(m.weight.shape, generator=g) + 0.5
            m.bias.data = torch.rand(m.bias.shape, generator=g) - 0.5
    return bb.cuda(), hd.cuda()


def oracle_run(bb, hd, img, gt):
    from oracle import stc_oracle as O
    bsd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in bb.state_dict().items()}
    hsd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in hd.state_dict().items()}
    nb, nh = {}, {}
    out = O.forward_train(bsd, hsd, img, gt, True, nb, nh)
    (out["loss_bce"] + out["loss_dice"]).backward()
    return out, bsd, hsd, nb, nh


def is_bn_cancelled_bias(name):
    """Conv biases feeding a train-mode BN: true gradient is 0 (SURVEY §7 traps), compare with an absolute floor."""
    return name.endswith((".conv.conv.0.bias", ".conv.conv.3.bias", "ca.conv1.bias")) or (".convs." in name and name.endswith(".0.bias"))


def run_case(stc, C, dtype, N, H, W, tol_fwd, tol_grad, ignore_border=True):
    bb, hd = build(stc, C, dtype)
    g = torch.Generator().manual_seed(123)
    img = torch.rand(N, 3, H, W, generator=g).cuda()
    gt = torch.randint(0, C, (N, 1, H, W), generator=g)
    if ignore_border:
        gt[:, :, :3] = 255
    gt = gt.cuda()
    ref, bsd, hsd, nb, nh = oracle_run(bb, hd, img, gt)
    feats = bb(img)
    losses = hd.forward_train(feats, None, gt, None)
    (losses["loss_bce"] + losses["loss_dice"]).backward()
    with torch.no_grad():
        bb.eval(); hd.eval()
    # forward quantities
    for k in ("loss_bce", "loss_dice"):
        assert abs(float(losses[k]) - float(ref[k])) <= tol_fwd * max(1.0, abs(float(ref[k]))), (k, float(losses[k]), float(ref[k]))
    assert abs(float(losses["acc_seg"]) - float(ref["acc_seg"])) < (0.05 if dtype == "fp32" else 2.0)
    # gradients
    worst = ("", 0.0)
    for mod, sd in ((bb, bsd), (hd, hsd)):
        for name, p in mod.named_parameters():
            gref = sd[name].grad
            assert p.grad is not None, name
            if is_bn_cancelled_bias(name):
                scale = float(gref.abs().max()) + 1e-3
                assert float((p.grad - gref).abs().max()) < (1e-3 if dtype == "fp32" else 5e-2) * max(1.0, scale), name
                continue
            r = rel_l2(p.grad, gref)
            if r > worst[1]:
                worst = (name, r)
            assert r <= tol_grad, (name, r)
    # running statistics after one step
    for mod, new in ((bb, nb), (hd, nh)):
        sd = mod.state_dict()
        for k, v in new.items():
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v), k
            else:
                assert rel_l2(sd[k], v) <= tol_fwd * 5, k
    return worst


def logits_case(stc, C, dtype, N, H, W):
    from oracle import stc_oracle as O
    bb, hd = build(stc, C, dtype)
    g = torch.Generator().manual_seed(7)
    img = torch.rand(N, 3, H, W, generator=g).cuda()
    bsd, hsd = bb.state_dict(), hd.state_dict()
    with torch.no_grad():
        ref = O.head_forward(hsd, O.backbone_forward(bsd, img, True, None), True, None)
        out = hd(bb(img))
    return out, ref


@pytest.mark.parametrize("stc", [False, True])
@pytest.mark.parametrize("C", [2, 3])
def test_fp32_parity_small(stc, C):
    out, ref = logits_case(stc, C, "fp32", 2, 64, 64)
    assert rel_l2(out, ref) <= 1e-4
    assert float((out.argmax(1) == ref.argmax(1)).float().mean()) >= 0.999
    run_case(stc, C, "fp32", 2, 64, 64, 1e-4, 1e-4)


@pytest.mark.parametrize("stc", [False, True])
def test_bf16_parity_small(stc):
    out, ref = logits_case(stc, 3, "bf16", 2, 64, 64)
    assert rel_l2(out, ref) <= 2e-2
    run_case(stc, 3, "bf16", 2, 64, 64, 2e-2, 6e-2)


def test_fp32_parity_config1_unet_512():
    """BASELINE.json configs[0]: U-Net fwd+bwd, batch 2 of 3x512x512, 3 classes, fp32 (oracle evaluated on the GPU)."""
    out, ref = logits_case(False, 3, "fp32", 2, 512, 512)
    assert rel_l2(out, ref) <= 1e-4
    assert float((out.argmax(1) == ref.argmax(1)).float().mean()) >= 0.999
    run_case(False, 3, "fp32", 2, 512, 512, 1e-4, 1e-4)


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
    bsd, hsd = bb.state_dict(), hd.state_dict()
    enc = lambda t: O.head_forward(hsd, O.backbone_forward(bsd, t, False, None), False, None)
    with torch.no_grad():
        ref_logits = O.slide_inference(enc, img, 3, (64, 64), (40, 40))
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
    import stc_unet_b200 as S
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
