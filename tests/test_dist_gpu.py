"""2-GPU data-parallel step (SyncBN statistics + gradient all-reduce through the NVLink peer-memory kernels) against the oracle on
the concatenated batch - BASELINE.json configs[2] semantics.  Needs two GPUs (`gpurun --gpus 2`); skipped on a single-GPU box.
The worker is tests/dist_worker.py, launched under torch.distributed.run."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(dtype, size, mode="posbn", env_extra=None, port=29531):
    env = dict(os.environ, **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), dtype, str(size), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    rep = [ln for ln in r.stdout.splitlines() if ln.startswith("DIST_REPORT ")]
    assert rep, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
    report = json.loads(rep[-1][len("DIST_REPORT "):])
    assert r.returncode == 0 and report["passed"], report
    return report


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")


@needs2
def test_two_rank_step_fp32_matches_oracle_on_concatenated_batch():
    rep = _run("fp32", 64)
    assert rep["exchange"] == "peer" and rep["grad_err_max_strict"] <= 1e-4 and rep["identical_grads_on_all_ranks"]


@needs2
def test_two_rank_step_bf16_256_halo_kernels():
    """bf16 at 256x256 per rank: W >= 128 puts levels 1-2 on the halo tcgen05 kernels with BN statistics from the conv epilogue."""
    rep = _run("bf16", 256, port=29532)
    assert rep["exchange"] == "peer" and rep["grad_err_median"] <= 0.1


@needs2
def test_two_rank_step_nccl_exchange_agrees():
    """STC_PEER=0: the same step with the NCCL exchanges (the fallback when symmetric memory is unavailable)."""
    rep = _run("fp32", 64, env_extra={"STC_PEER": "0"}, port=29533)
    assert rep["exchange"] == "nccl" and rep["grad_err_max_strict"] <= 1e-4
