"""The single-pass CSF step (k_csf_staged / k_csf_fused) against the three-pass step (LBM_CSF_FUSED=0) — helper of tests/test_gpu_csf.py.

Both steps state the same arithmetic in two differently shaped kernels.  With floating-point contraction off they agree
bit for bit (hardware: liblbm_b200_nofma.so, `make NOFMA=1`; the emulated device never contracts); in the product build
nvcc picks the multiply-adds it fuses per kernel, and the two differ in the last bit from the second step on (measured on
the B200: 1e-16 relative, growing to 7e-16 over nine steps; the same kernel against itself — run twice, other band heights,
the software-pipelined instantiation — stays bit-identical, so there is no race behind it).

As a script:  python tests/csf_fused_check.py --lib <path or suffix> R C rpb kernel   -> exit 0 iff bit-identical
TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"))

SCHEDULE = (1, 1, 3, 4)  # a first step out of an import (three passes either way), then fused steps with getters in between
NAMES = ("f_red", "f_blue", "Fs", "phase")


# the single-pass kernels: name -> (LBM_CSF_STAGED, LBM_CSF_PIPE)
KERNELS = {"staged+stash": ("2", "0"),   # k_csf_staged<3, STASH>: bulk-async staging + tensor-memory stash (the default)
           "staged": ("1", "0"),         # k_csf_staged<7, false>: rows resident in the stage slots
           "fused": ("0", "0"),          # k_csf_fused<false>: two pulls through registers
           "fused+pipe": ("0", "1")}     # k_csf_fused<true>: software-pipelined


def run(R, C, fused, pipe="0", rpb=0, schedule=SCHEDULE, kernel=None):
    """[(f_red, f_blue, Fs, phase)] after every entry of the schedule; the switches are read when the domain is created.
    kernel: one of KERNELS (overrides pipe); None leaves LBM_CSF_STAGED at the library's default"""
    import cases
    from oracle_lib import Oracle
    from test_gpu_csf import csf_params

    old = {k: os.environ.get(k) for k in ("LBM_CSF_FUSED", "LBM_CSF_PIPE", "LBM_TP_RPB", "LBM_CSF_STAGED")}
    if kernel is not None:
        os.environ["LBM_CSF_STAGED"], pipe = KERNELS[kernel]
    os.environ["LBM_CSF_FUSED"], os.environ["LBM_CSF_PIPE"] = str(fused), str(pipe)
    if rpb:
        os.environ["LBM_TP_RPB"] = str(rpb)
    else:
        os.environ.pop("LBM_TP_RPB", None)
    try:
        st = Oracle().csf_init(csf_params(R, C))
        d = cases.csf(R, C)
        d.init_two_phase(st["r_rho"], st["b_rho"], st["u"])
        out = []
        for n in schedule:
            d.step(n)
            out.append((d.get_f(0), d.get_f(1), d.get_interfacial_tension(), d.get_phase()[0]))
        d.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out


def first_difference(A, B):
    """None when every field of every snapshot is bit-identical, else a description of the first one that is not"""
    for n, (a, b) in enumerate(zip(A, B)):
        for name, x, y in zip(NAMES, a, b):
            if not np.array_equal(x, y):
                k = tuple(int(v) for v in np.argwhere(x != y)[0])
                return f"snapshot {n}, {name}: {int((x != y).sum())} entries differ, first at {k}, max |delta| {np.abs(x - y).max():.3e}"
    return None


def worst_relative(A, B):
    w = 0.0
    for a, b in zip(A, B):
        for x, y in zip(a[:2], b[:2]):
            w = max(w, float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)))
    return w


if __name__ == "__main__":
    import lbm_b200 as L

    argv = sys.argv[1:]
    if argv[:1] == ["--lib"]:
        L.LIB_PATH = argv[1] if "/" in argv[1] else os.path.join(L.PKG_DIR, f"liblbm_b200_{argv[1]}.so")
        argv = argv[2:]
    R, C, rpb, kernel = int(argv[0]), int(argv[1]), int(argv[2]), argv[3]
    diff = first_difference(run(R, C, 1, rpb=rpb, kernel=kernel), run(R, C, 0, rpb=rpb))
    print(f"{os.path.basename(L.LIB_PATH)} {R}x{C} rpb={rpb} {kernel}: " + (diff or "single pass == three passes, bit for bit"))
    sys.exit(1 if diff else 0)
