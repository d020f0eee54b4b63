"""Multi-GPU parity as a driver-visible test: tests/multigpu_check.py under torch.distributed.run on 2, 4 and 8 GPUs of
the box (skipped where the box has fewer).  The sharded chains (NMF and tri-factorisation; Gibbs, ICM, VB, NP; small
shapes and a K = 50 strip of the benchmark matrix in the benchmark's packed layout) must be BITWISE the single-GPU
chains, and the class API must give every rank rank 0's chain without a common numpy seed."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_chains_are_bitwise_the_single_gpu_chains(n):
    if _gpus() < n:
        pytest.skip("needs %d GPUs, the box has %d" % (n, _gpus()))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    out = res.stdout + res.stderr
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        with open(os.path.join(log, "multigpu_check_n%d.log" % n), "w") as fh:
            fh.write(out)
    assert res.returncode == 0 and "MULTIGPU_CHECK PASS" in out, out[-4000:]
    assert "MISMATCH" not in out
