"""Decomposed run on >= 2 GPUs (one z-slab per rank, halo planes and Krylov sums over NCCL) against the CPU oracle's
emulation of the same ranks.  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["implicit", "explicit", "bicg", "backward", "cn"])
def test_two_rank_parity(mode):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "run_multi.py"), mode, "30"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_two_rank_function_objects():
    """meltPoolDimensions and solidificationData across the slab interface (processor faces on every cell)."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tools", "run_multi.py"), "funcobj", "120"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_two_rank_random_cases():
    """The seeded random cases of test_gpu_fuzz.py on two z-slabs against the oracle's emulation of two ranks."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(ROOT, "tools", "run_multi.py"), "fuzz", "16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_two_rank_ldu_solver_with_processor_interfaces():
    """b200_ldu_* with coupled interfaces (the decomposed bPCG seam): Amul and none/diagonal-preconditioned solves reproduce the
    oracle's serial solve of the global matrix, block-local DIC/DILU solve the global system."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(ROOT, "tools", "run_multi_ldu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTILDU ALL OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
