"""Two ranks on two GPUs (skipped on a one-GPU box): the reference's own wrapping -- convert_sync_batchnorm +
DistributedDataParallel (tools/train.py:216-229) -- over the mirror, one sample per rank, must reproduce the reference's
single-process B=2 golden step: G losses / predictions / averaged gradient norms / synced running statistics, and the
stacked, eagerly evaluated D step.  The body is tests/dist_syncbn_check.py (run under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [["fp32"], ["fp32", "graphs"], ["fp32", "eager", "skip"], ["bf16", "graphs", "skip"],
                                  ["fp32", "eager", "p2p"], ["fp32", "graphs", "skip", "p2p"], ["bf16", "graphs", "skip", "p2p"]],
                         ids=["fp32", "fp32-graphs", "fp32-skipdead", "bf16-graphs-skipdead", "fp32-p2p", "fp32-graphs-skipdead-p2p",
                              "bf16-graphs-skipdead-p2p"])
def test_ddp_syncbn_two_ranks_match_single_process_reference(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + (os.getpid() + len(" ".join(mode))) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_syncbn_check.py")] + mode
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "DIST_SYNCBN_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
