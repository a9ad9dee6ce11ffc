"""CPU: the product's collision arithmetic compiled for the host (tests/host_kernels/host_kernels.cu: the very
__host__ __device__ functions of csrc/lbm_device.cuh that the kernels inline) against the oracle.  Lets a change of the
arithmetic be checked here, where there is no GPU; the GPU suite then only has to confirm the fused-multiply-add build."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import cases
from oracle_lib import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_kernels", "host_kernels.cu")
OUT = os.path.join(HERE, "host_kernels", "_build", "libhost_kernels.so")
DEPS = [SRC, os.path.join(os.path.dirname(HERE), "lattice-boltzmann-method_b200", "csrc", "lbm_device.cuh")]
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

CXI = (0, 1, 0, -1, 0, 1, -1, -1, 1)
CYI = (0, 0, 1, 0, -1, 1, 1, -1, -1)


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    dp = C.POINTER(C.c_double)
    lib.host_kbc_collide.argtypes = [dp, dp, dp, C.c_int, C.c_long, C.c_double]
    return lib


def ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def advect(f):
    """solver::advect / kbc::advect: periodic streaming"""
    out = np.empty_like(f)
    for q in range(9):
        out[..., q] = np.roll(f[..., q], (CXI[q], CYI[q]), axis=(0, 1))
    return out


@pytest.mark.parametrize("given", [False, True])
def test_kbc_collision_on_the_host_matches_the_oracle(host, given):
    """double shear layer (test/ulbm_double_shear_flow.cpp) with noise on the populations, 12 steps: collision from the
    product's source, streaming in numpy, against orc_kbc_step; `given`: the caller's m0, u differ from the populations'"""
    orc = Oracle()
    R = 48
    s2 = 1.0 / (0.5 + 3.0 * 1.70766666e-4)
    m0, u = cases.double_shear_fields(R, R)
    f = orc.kbc_equilibrium(m0, u, fresh_object=True)
    f *= 1.0 + 1e-3 * np.random.default_rng(3).standard_normal(f.shape)
    if given:
        m0 = m0 * (1.0 + 1e-4 * np.random.default_rng(4).standard_normal(m0.shape))
        u = u + 1e-5 * np.random.default_rng(5).standard_normal(u.shape)
    else:
        m0 = f.sum(-1)
        u = np.stack([(f * np.array(CXI)).sum(-1), (f * np.array(CYI)).sum(-1)], -1) / m0[..., None]
    ref_f, ref_m0, ref_u = f.copy(), np.ascontiguousarray(m0), np.ascontiguousarray(u)
    mine = f.copy()
    mine_m0, mine_u = ref_m0.copy(), ref_u.copy()
    for step in range(12):
        orc.kbc_step(ref_f, ref_m0, ref_u, s2)
        c = np.ascontiguousarray(mine)
        host.host_kbc_collide(ptr(c), ptr(mine_m0), ptr(mine_u), 1 if (given and step == 0) else 0, c.size // 9, s2)
        mine = advect(c)
        assert cases.relerr(mine, ref_f) < 1e-13, step
