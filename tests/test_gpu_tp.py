"""Tensor-parallel decode on 2 GPUs vs the single-GPU engine (needs >= 2 GPUs; skipped on a one-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="tensor parallelism needs at least two GPUs")
@pytest.mark.parametrize("engine", ["fused", "fused-p2p", "fused-barrier", "nccl"])
def test_tp2_tokens_equal_single_gpu(engine):
    """fused: the persistent kernel with peer stores over NVLink, the partials exchanged point to point per column slice as tagged
    8-byte words the reader polls (default, "ll");
    fused-p2p: the same exchange with separate flags behind a system-scope release;
    fused-barrier: a barrier across the GPUs + a reduce phase instead (round 1's default);
    nccl: the per-op engine with ncclAllReduce after every row-parallel GEMV (the baseline of SURVEY.md 8e)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", {"fused": "29533", "fused-p2p": "29535", "fused-barrier": "29536", "nccl": "29534"}[engine],
           os.path.join(ROOT, "scripts", "tp_check.py")]
    env = dict(os.environ)
    env.pop("TURBOINFER_B200_TP_ENGINE", None)
    env.pop("TURBOINFER_B200_TP_REDUCE", None)
    if engine == "nccl":
        env["TURBOINFER_B200_TP_ENGINE"] = "nccl"
    if engine == "fused-p2p":
        env["TURBOINFER_B200_TP_REDUCE"] = "p2p"
    if engine == "fused-barrier":
        env["TURBOINFER_B200_TP_REDUCE"] = "barrier"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0 and "TP CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
