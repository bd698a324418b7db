"""Tensor-parallel numerics ON HARDWARE: TP-2 logits == single-GPU logits of the same model, for every implementation of
the exchange step (gpt-fast/tp.py:120-121,139-140): fused into the row-parallel GEMV's finalisers, the one-shot
peer-memory kernel, ncclAllReduce.  Needs >= 2 GPUs on the box (skipped otherwise; the world-size-2 gloo tests of
tests/test_tp_gloo.py cover the host logic on CPU).  Kept runs: profiles/r02_tp*_check_*.jsonl."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["gemv", "fused", "nccl"])
def test_tp2_logits_equal_single_gpu(exchange, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, TEAL_TP_ALLREDUCE=exchange, MASTER_ADDR="127.0.0.1")
    out = tmp_path / "tp.jsonl"
    port = 29540 + ["gemv", "fused", "nccl"].index(exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "tp_check.py"), "--out", str(out)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "TP_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
