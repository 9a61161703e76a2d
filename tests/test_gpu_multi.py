"""GPU, >= 2 devices: eigen_s on a 2D cyclic grid (NCCL) vs the oracle.  Skipped on one GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nproc,n,mtype,mode,solver", [(2, 600, 2, "A", "s"), (2, 257, 0, "A", "s"), (2, 500, 2, "N", "s"),
                                                       (2, 1300, 2, "A", "s"), (4, 700, 2, "A", "s"), (4, 1500, 0, "A", "s"),
                                                       (8, 900, 2, "A", "s"),
                                                       # penta-diagonal driver on the same grids
                                                       (2, 601, 2, "A", "sx"), (2, 1300, 0, "A", "sx"), (2, 500, 2, "N", "sx"),
                                                       (4, 1500, 2, "A", "sx")])
# (eigen_sx on the 2x4 grid: not yet run on hardware in round 1 -- its grid-dependent code, symv_strip / pvec partial sums /
#  staircase GEMM, is the code eigen_s exercises on 2x4; add the case once an 8-GPU slot has confirmed it)
def test_eigen_s_multi_rank(nproc, n, mtype, mode, solver):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + nproc), os.path.join(ROOT, "tools", "run_multi.py"), str(n),
           str(mtype), mode, "check", solver]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["w_err_over_tol"] <= 1.0
    if mode != "N":
        assert out["residual"] <= 10 and out["orth"] <= 10
