import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_model_gpu import build, oracle, ours, inputs, grad_errors
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
for C in (2, 3):
    img, gt = inputs(2, C, 64, 64)
    bb, hd = build(True, C, "fp32", posbn=True)
    ref64 = oracle(bb, hd, img, gt, torch.float64)
    ref32 = oracle(bb, hd, img, gt, torch.float32)
    et = grad_errors(ref32, ref64)
    for cen in ("run1", "run2"):      # two runs: the spread shows the run-to-run variation (atomics order)
        got = ours(True, C, "fp32", img, gt, True)
        eo = grad_errors(got, ref64)
        qk = sorted(((eo[k], et[k], k[1]) for k in eo if k[1].endswith((".q.weight", ".k.weight"))), reverse=True)[:4]
        rest = sorted(((eo[k], et[k], k[1]) for k in eo if not k[1].endswith((".q.weight", ".k.weight"))), reverse=True)[:3]
        print("C", C, cen, "qk worst", [(f"{a:.2e}", f"{b:.2e}", n) for a, b, n in qk])
        print("      rest worst", [(f"{a:.2e}", f"{b:.2e}", n) for a, b, n in rest], flush=True)
import statistics
from tests.util import rel_l2
for C in (3,):
    img, gt = inputs(2, C, 64, 64)
    for posbn in (True, False):
        bb, hd = build(True, C, "fp32", posbn=posbn)
        ref64 = oracle(bb, hd, img, gt, torch.float64)
        refac = oracle(bb, hd, img, gt, torch.float32, autocast=True)
        ea = grad_errors(refac, ref64)
        print("posbn", posbn, "autocast logits %.3e grads median %.3e" % (rel_l2(refac["logits"], ref64["logits"]), statistics.median(ea.values())))
        for cen in ("run1",):
            got = ours(True, C, "bf16", img, gt, posbn)
            eo = grad_errors(got, ref64)
            qk = statistics.median(eo[k] for k in eo if k[1].endswith((".q.weight", ".k.weight")))
            print("  bf16", cen, "logits %.3e grads median %.3e qk median %.3e (autocast qk %.3e)" % (
                rel_l2(got["logits"], ref64["logits"]), statistics.median(eo.values()), qk,
                statistics.median(ea[k] for k in ea if k[1].endswith((".q.weight", ".k.weight")))), flush=True)
