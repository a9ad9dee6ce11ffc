"""CPU: register / local-memory budget of the step kernels as compiled into liblbm_b200.so (cuobjdump -res-usage).
The hot kernels are occupancy-sensitive: a stray launch-bounds argument once grew the BGK+IBM kernel from 76 to 126
registers and cost the headline workload 8 % without any test noticing.  Budgets = what the measured builds use."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lattice-boltzmann-method_b200", "liblbm_b200.so")

# mangled-name fragment -> (max registers, what it is)
BUDGET = {
    "k_bgk_interiorILi1ELi0ELi0ELb0E": (80, "BGK pull, compressible, no force (cylinder_bb)"),
    "k_bgk_interiorILi1ELi0ELi2ELb0E": (80, "BGK pull, compressible, IBM force field (cylinder, the headline)"),
    "k_bgk_interiorILi1ELi1ELi0ELb0E": (80, "BGK pull, incompressible (Poiseuille)"),
    "k_bgk_interiorILi1ELi0ELi0ELb1E": (100, "BGK + advection-diffusion lattice (sedimentation)"),
    "k_bgk_interiorILi1ELi2ELi0ELb0E": (128, "KBC: 4 resident blocks of 128 threads"),
    "k_tp_stagedILi0ELi3ELi3ELb1E": (168, "MRT colour gradient, staged rows + tensor-memory stash (the default): 3 resident blocks"),
    "k_tp_stagedILi1ELi4ELi2ELb1E": (168, "Rothman-Keller, staged rows + tensor-memory stash (the default): 2 resident blocks by shared memory"),
    "k_tp_stagedILi0ELi5ELi2ELb0E": (255, "MRT colour gradient, rows resident in the stage slots (LBM_TP_STASH=0): 2 resident blocks"),
    "k_csf_stagedILi3ELb1E": (208, "CSF single pass, staged rows + tensor-memory stash (the default): 2 resident blocks"),
    "k_tp_fusedILi0ELb1E": (168, "MRT colour gradient, pipelined: 3 resident blocks"),
    "k_tp_fusedILi1ELb1E": (168, "Rothman-Keller, pipelined: 3 resident blocks"),
    "k_csf_collide_ringILi1E": (168, "CSF collision pass: 3 resident blocks"),
    "k_csf_fusedILb0E": (168, "CSF single pass, plain: 3 resident blocks"),
    "k_csf_fusedILb1E": (255, "CSF single pass, pipelined: 2 resident blocks"),
    "k_bgk_interiorILi1ELi0ELi2ELb1E": (110, "BGK + advection-diffusion lattice + IBM force field (sedimentation_ibm)"),
}


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_step_kernels_stay_within_their_register_budget():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    found = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+?):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:\d+ LOCAL:(\d+)", line)
        if m and name:
            for frag in BUDGET:
                if frag in name:
                    found[frag] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
            name = None
    for frag, (limit, what) in BUDGET.items():
        assert frag in found, f"{what}: kernel {frag} not found in the library"
        reg, stack, local = found[frag]
        assert reg <= limit, f"{what}: {reg} registers > {limit}"
        assert stack == 0 and local == 0, f"{what}: spills (stack {stack}, local {local})"
