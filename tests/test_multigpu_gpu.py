"""GPU test of the multi-GPU path: needs >= 2 CUDA devices (skipped otherwise).  Launches tools/multigpu_check.py with
torchrun: a dam-break block split into two x-slabs, moving towards +x so that particles migrate, must reproduce the
single-GPU run of the same scene (identical iteration counts, fields by particle id)."""
import os
import subprocess
import sys

import pytest

from tests.parity import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("prec,axis,p2p,host", [("f64", 0, "1", 0), ("f32", 2, "1", 0), ("f64", 2, "0", 0), ("f64", 0, "1", 1)])
def test_two_slabs_reproduce_single_gpu(prec, axis, p2p, host):
    """p2p = "1": NVLink peer-memory refresh + fused all-reduce; "0": NCCL send/recv + ncclAllReduce.
    host = 1: every slab step goes through dfsph_b200_step_host (host buffers in device order, migration inside).
    The double runs go 25 free-running steps, through the impact of the moving block on the wall; the float run stops at
    12 steps, before the impact: collisions amplify float rounding differences (the neighbour order of a slab run differs
    from the single-GPU run's, so sums round differently) from 3e-6 at step 12 to 3e-3 at step 25, identically with the
    NVLink and the NCCL transport (profiles/r1_multigpu_f32_drift.md)."""
    steps = "25" if prec == "f64" else "12"
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "multigpu_check.py"), prec, "small", steps, str(axis), str(host)]
    os.environ["DFSPH_B200_P2P"] = p2p
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
