"""CPU-only checks of the drop-in boundary: the C-ABI library builds/loads and exports every declared symbol, the
module mirror has the reference's constructor surface and state_dict layout, and nothing silently runs on the CPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from stc_unet_b200 import build
    from stc_unet_b200._lib import HEADER_PATH, lib, parse_header
    path = build.build()
    assert os.path.exists(path)
    protos = parse_header(HEADER_PATH)
    declared = set(re.findall(r"\b(stc_\w+)\s*\(", re.sub(r"/\*.*?\*/", " ", open(HEADER_PATH).read(), flags=re.S)))
    assert declared == set(protos), declared ^ set(protos)
    dll = ctypes.CDLL(path)
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/stc_b200.h but not exported"
    assert lib.raw("stc_version")() >= 100
    assert len(protos) >= 50


def test_no_cpu_fallback():
    import stc_unet_b200 as S
    bb = S.build_backbone(dict(type="UnetBackbone", in_channels=3, channel_list=[64, 128, 256, 512]))
    with pytest.raises(RuntimeError, match="no CPU"):
        bb(torch.rand(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU"):
        S.ops.seg_loss(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            S._lib.lib.ensure_device(0)


def test_registry_contract_and_constructor_surface():
    import stc_unet_b200 as S
    for name in ("UnetBackbone", "UnetHead", "CrossEntropyLoss", "DiceLoss", "EncoderDecoder"):
        assert S.MODELS.get(name) is not None or S.MODELS.get(name + "B200") is not None
    with pytest.raises(KeyError):
        S.build_backbone(dict(type="NoSuchBackbone"))
    with pytest.raises(ValueError):
        S.build_head(dict(type="UnetHead", num_classes=3, out_channels=2))
    with pytest.warns(UserWarning, match="binary segmentation"):
        hd = S.build_head(dict(type="UnetHead", num_classes=2, loss_decode=dict(type="CrossEntropyLoss", loss_name="loss_ce")))
    assert hd.out_channels == 2 and hd.num_classes == 2 and hd.align_corners is False and hd.ignore_index == 255
    assert hd.dropout is not None and hd.conv_seg.weight.shape == (2, 64, 1, 1)
    hd.init_weights()
    assert float(hd.conv_seg.weight.std()) < 0.05 and float(hd.conv_seg.bias.abs().max()) == 0.0
    with pytest.warns(UserWarning):
        hd1 = S.build_head(dict(type="UnetHead", num_classes=2, out_channels=1))
    assert hd1.threshold == 0.3
    seg = S.build_segmentor(dict(type="EncoderDecoder", backbone=dict(type="UnetBackbone"), decode_head=dict(type="UnetHead", num_classes=3)),
                            test_cfg=dict(mode="whole"))
    keys = list(seg.state_dict())
    assert keys[0] == "backbone.inc.conv.conv.0.weight" and "decode_head.conv_seg.weight" in keys
    assert isinstance(seg.backbone.inc.conv.conv[1], torch.nn.SyncBatchNorm)


def test_grad_arena_layout_and_bucket_plan():
    from stc_unet_b200.ops import GradArena
    from stc_unet_b200.train import plan_buckets
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 64, 3, 1000, 7)]
    arena = GradArena(ps)
    offs = [arena.offsets[id(p)][0] for p in ps]
    assert offs == [0, 8, 72, 76, 1076] and arena.total == 1084 and all(o % 4 == 0 for o in offs)
    v = arena.view(ps[3])
    assert v.shape == (1000,) and v.data_ptr() == arena.flat.data_ptr() + 76 * 4
    buckets, owner = plan_buckets([(o, p.numel()) for o, p in zip(offs, ps)], arena.total, cap_elems=1000)
    # buckets are contiguous, cover the arena, and are ordered from the END (backward order)
    assert buckets[0][1] == arena.total and buckets[-1][0] == 0
    assert all(buckets[i][0] == buckets[i + 1][1] for i in range(len(buckets) - 1))
    assert sum(b[2] for b in buckets) == len(ps) and owner[4] == 0 and owner[0] == len(buckets) - 1


def test_metrics_from_confusion_follow_reference_definitions():
    """tests/test_metrics.py:29-85 (legacy_mean_iou / legacy_mean_dice / legacy_mean_fscore from the bincount matrix), incl.
    nan_to_num; and NOT the fork's inflated values (metrics.py:454-457)."""
    import numpy as np
    from oracle import stc_oracle as O
    from stc_unet_b200.metrics import metrics_from_confusion
    rng = np.random.RandomState(0)
    C = 19
    pred = rng.randint(0, C, size=(10, 30, 30)); label = rng.randint(0, C, size=(10, 30, 30)); label[:, 2, 5:10] = 255
    label[label == 7] = 3        # an absent class -> NaNs
    pred[pred == 7] = 3
    cm = O.confusion_matrix(pred, label, C, 255)
    tot = cm.astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        all_acc = np.diag(tot).sum() / tot.sum()
        acc = np.diag(tot) / tot.sum(axis=1)
        iou = np.diag(tot) / (tot.sum(axis=1) + tot.sum(axis=0) - np.diag(tot))
        dice = 2 * np.diag(tot) / (tot.sum(axis=1) + tot.sum(axis=0))
        prec = np.diag(tot) / tot.sum(axis=0)
        fscore = 2 * prec * acc / (prec + acc)
    m = metrics_from_confusion(cm, ["mIoU", "mDice", "mFscore"])
    assert np.allclose(m["aAcc"], all_acc) and np.allclose(m["Acc"], acc, equal_nan=True) and np.allclose(m["IoU"], iou, equal_nan=True)
    assert np.allclose(m["Dice"], dice, equal_nan=True) and np.allclose(m["Fscore"], fscore, equal_nan=True)
    assert np.allclose(m["Precision"], prec, equal_nan=True) and np.allclose(m["Recall"], acc, equal_nan=True)
    assert np.isnan(m["IoU"][7])
    m2 = metrics_from_confusion(cm, "mIoU", nan_to_num=-1)
    assert m2["IoU"][7] == -1 and m2["Acc"][7] == -1
    with pytest.raises(KeyError):
        metrics_from_confusion(cm, ["unsupported"])
    # uniformly random 3-class predictions score ~1/3 accuracy (the fork's tampered code reports 0.80 here)
    p3, l3 = rng.randint(0, 3, 10000), rng.randint(0, 3, 10000)
    assert abs(metrics_from_confusion(O.confusion_matrix(p3, l3, 3))["aAcc"] - 1 / 3) < 0.03


def test_poly_lr_matches_the_reference_schedule():
    """lr_config = dict(policy='poly', power=0.9, min_lr=1e-6, by_epoch=True), 50 epochs, Adam lr 1e-5 (my_config/STC-UNet.py:87-93):
    mmcv's PolyLrUpdaterHook formula, applied through param_groups like the hook does."""
    import stc_unet_b200 as S
    p = torch.nn.Parameter(torch.zeros(3))
    opt = S.build_optimizer(torch.nn.ParameterList([p]), dict(type="Adam", lr=1e-5, betas=(0.9, 0.999)))
    assert isinstance(opt, torch.optim.Adam)
    upd = S.PolyLrUpdater(opt, max_progress=50, power=0.9, min_lr=1e-6, by_epoch=True)
    seen = []
    for epoch in range(50):
        upd.before_epoch(epoch)
        upd.before_iter(123)          # ignored: by_epoch
        seen.append(opt.param_groups[0]["lr"])
    assert seen[0] == pytest.approx(1e-5)
    assert seen[25] == pytest.approx((1e-5 - 1e-6) * 0.5 ** 0.9 + 1e-6)
    assert seen[49] == pytest.approx((1e-5 - 1e-6) * (1 / 50) ** 0.9 + 1e-6)
    assert all(a > b for a, b in zip(seen, seen[1:]))
    assert S.poly_lr(0.01, 40000, 40000, 1.0, 1e-4) == pytest.approx(1e-4)
    with pytest.raises(KeyError):
        S.build_optimizer(torch.nn.ParameterList([p]), dict(type="NoSuchOptimizer", lr=1.0))


def test_checkpoint_round_trip_uses_the_reference_layout(tmp_path):
    """save_checkpoint / load_checkpoint: mmcv's {'meta', 'state_dict'} layout, the reference's keys (backbone.* / decode_head.*), fp32
    tensors on the CPU, nothing derived (packed bf16 operands) serialised; loading copies in place; 'module.' prefixes are stripped."""
    import stc_unet_b200 as S
    torch.manual_seed(0)
    cfg_b = dict(type="UnetBackbone", in_channels=3, channel_list=[64, 128, 256, 512], context_layer="kernelselect", transformer_block=True)
    cfg_h = dict(type="UnetHead", se=True, num_classes=3, channels=64, threshold=0.2, norm_cfg=dict(type="BN", requires_grad=True))
    seg = S.EncoderDecoder(cfg_b, cfg_h)
    path = str(tmp_path / "epoch_1.pth")
    S.save_checkpoint(seg, path, meta=dict(epoch=1, iter=10))
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"meta", "state_dict"} and ck["meta"]["epoch"] == 1
    sd = ck["state_dict"]
    assert len(sd) == 233 + 102 and all(k.startswith(("backbone.", "decode_head.")) for k in sd)
    assert all(v.device.type == "cpu" and v.dtype in (torch.float32, torch.int64) for v in sd.values())
    seg2 = S.EncoderDecoder(cfg_b, cfg_h)
    ptrs = {k: v.data_ptr() for k, v in seg2.state_dict().items()}
    wrapped = OrderedDictPrefix(sd, "module.")
    torch.save(dict(state_dict=wrapped), path)
    S.load_checkpoint(seg2, path, strict=True)
    for k, v in seg2.state_dict().items():
        assert v.data_ptr() == ptrs[k]            # in place
        assert torch.equal(v, sd[k]), k
    torch.save(dict(state_dict={"backbone.bogus": torch.zeros(1)}), path)
    with pytest.raises(RuntimeError):
        S.load_checkpoint(seg2, path, strict=True)


def OrderedDictPrefix(sd, prefix):
    from collections import OrderedDict
    return OrderedDict((prefix + k, v) for k, v in sd.items())


def test_step_cache_record_finalize_layout_on_cpu():
    """StepCache host logic without a GPU: a recorded step's packs get 16-byte aligned slots in ONE arena in first-use order (so q/k/v packs
    are equally spaced - what the batched folded-weight products rely on), the batched-pack table carries {src ptr, dst offset, shape, mode},
    work-item prefix sums are dense, and wgrad workspaces are handed out in recorded order from one zero-able arena."""
    from stc_unet_b200 import ops
    c = ops.StepCache()
    c.begin_step()
    assert c.mode == "record"
    ws_ = [torch.randn(512, 512, 1, 1) for _ in range(3)] + [torch.randn(64, 3, 3, 3), torch.randn(30, 64, 3, 3)]
    shapes = []
    for w in ws_[:3]:
        assert c.pack(w, torch.bfloat16, 0, 512, (1, 512, 512)) is None
        shapes.append((1, 512, 512))
    assert c.pack(ws_[3], torch.bfloat16, 2, 64, (1, 64, 64)) is None            # im2col pack of the image conv
    assert c.pack(ws_[4], torch.bfloat16, 1, 30, (9, 64, 30)) is None            # dgrad orientation, odd Cout
    assert c.pack(ws_[0], torch.bfloat16, 0, 512, (1, 512, 512)) is None and len(c.order) == 5   # asked twice, recorded once
    for n in (9 * 64 * 30, 512 * 1536, 7):
        assert c.workspace(n, torch.device("cpu")).numel() == n
    c.mode = "record"
    c._finalize()
    table, prefix = c.table.tolist(), c.prefix.tolist()
    assert [r[0] for r in table] == [w.data_ptr() for w in ws_]
    offs = [r[1] for r in table]
    assert all(o % 8 == 0 for o in offs) and offs[1] - offs[0] == offs[2] - offs[1] == 512 * 512      # 16 B aligned, q/k/v equally spaced
    assert [r[7] for r in table] == [0, 0, 0, 2, 1]
    assert prefix == [0, 262144, 524288, 786432, 786432 + 64 * 64, 786432 + 64 * 64 + 30 * 64] and c.total == prefix[-1]
    assert c.ws_off == [0, 9 * 64 * 30, 9 * 64 * 30 + 512 * 1536, 9 * 64 * 30 + 512 * 1536 + 8]
    assert c.arena.numel() >= offs[-1] + 9 * 64 * 30 and c.ws_arena.numel() == c.ws_off[-1]


def test_widen_rule_and_crop_flip_draws():
    """Pure host logic: (1) which narrow bf16 convs are zero-padded onto the tensor-core kernels (ops._widen), (2) draw_crop_flip consumes the
    random stream exactly like the reference pipeline does per image - RandomCrop.get_crop_bbox (transforms.py:599-608: randint(0, margin_h+1),
    randint(0, margin_w+1)) followed by RandomFlip (np.random.rand() < prob)."""
    import numpy as np
    from stc_unet_b200 import ops
    bf, G = torch.bfloat16, 2e9
    assert ops._widen(16, 16, bf, False, G) == (64, 32)            # UNet++ x_0_4: K chunk of 64 input channels, N tile of 32
    assert ops._widen(576, 32, bf, True, G) == (576, 64)           # wgrad needs 64-multiples on both sides
    assert ops._widen(32, 576, bf, False, G) == (64, 576)
    assert ops._widen(64, 64, bf, False, G) is None                # already eligible
    assert ops._widen(16, 16, bf, False, 1e8) is None              # tiny layers (CoordAtt) stay on the SIMT engine
    assert ops._widen(16, 16, torch.float32, False, G) is None     # fp32 parity path is never touched
    assert ops._widen(8, 64, bf, False, G) is None                 # > 4x padding
    assert ops._widen(20, 64, bf, False, G) is None                # not a multiple of 8
    rs_a, rs_b = np.random.RandomState(7), np.random.RandomState(7)
    g = ops.draw_crop_flip(5, (600, 600), (512, 512), 0.5, rs_a).numpy()
    for i in range(5):
        assert g[i, 0] == rs_b.randint(0, 89) and g[i, 1] == rs_b.randint(0, 89) and g[i, 2] == int(rs_b.rand() < 0.5)
    assert g[:, :2].min() >= 0 and g[:, :2].max() <= 88
    small = ops.draw_crop_flip(4, (40, 70), (64, 64), 1.0, np.random.RandomState(1)).numpy()
    assert (small[:, 0] == 0).all() and (small[:, 1] <= 6).all() and (small[:, 2] == 1).all()      # no vertical margin; always flipped
