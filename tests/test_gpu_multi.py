"""GPU (>= 2 devices): one process per GPU over NCCL — tests/mp_nccl_check.py under torchrun.  Poiseuille
(pressure packets across the ring), cylinder (IBM, also with the body across the cut) and the CSF model must equal
the monolithic run bit for bit; MRTCG and RK (moment-plane halos) must stay on the oracle to 1e-12."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_devices():
    import torch

    return torch.cuda.device_count()


@pytest.mark.skipif("n_devices() < 2")
def test_nccl_slab_ring_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join("tests", "mp_nccl_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for case in ("poiseuille", "mrtcg", "rk", "csf", "cylinder"):
        assert f"{case} ring of 2" in r.stdout
    assert "cylinder across the cuts, ring of 2: bit-exact vs monolithic = True" in r.stdout


@pytest.mark.skipif("n_devices() < 4")
def test_blocks_bound_across_four_ranks():
    """column faces bound across ranks (NCCL send / recv of the face tails): bit-identical to the linked blocks of one process"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
           "--master-port", "29518", os.path.join(ROOT, "tests", "mp_blocks_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "bit-exact vs linked blocks after 65 steps = True" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
