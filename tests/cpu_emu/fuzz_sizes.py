"""TEST INFRASTRUCTURE (see cuda_emu.hpp): the size-parametrised GPU parity tests called with RANDOM grid sizes on the
emulated device (run it under the AddressSanitizer build to also catch out-of-bounds accesses at odd sizes):
    make -C tests/cpu_emu SAN=1
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 \
        LBM_EMU=1 LBM_EMU_ASAN=1 python tests/cpu_emu/fuzz_sizes.py [seed [rounds]]
Not part of the default suite (minutes); a failing size is printed so that it can be added to the tests' own lists."""
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "lattice-boltzmann-method_b200", "python")]
os.environ.setdefault("LBM_EMU", "1")
import conftest  # noqa: E402,F401  (points the binding at the emulated build)

import numpy as np  # noqa: E402

import lbm_b200 as L  # noqa: E402
import test_gpu_bgk as bgk  # noqa: E402
import test_gpu_csf as csf  # noqa: E402
import test_gpu_kbc as kbc  # noqa: E402
import test_gpu_two_phase as tp  # noqa: E402
from oracle_lib import Oracle  # noqa: E402


class EnvPatch:
    """the part of pytest's monkeypatch the single-pass test uses"""

    def setenv(self, k, v):
        os.environ[k] = v


def main(seed, rounds):
    rng = np.random.default_rng(seed)
    orc = Oracle()
    pick = lambda lo, hi: int(rng.integers(lo, hi + 1))  # noqa: E731
    failed = []
    for i in range(rounds):
        cases = [
            ("streaming", lambda X, Y: bgk.test_streaming_is_bit_exact(orc, X, Y), (pick(3, 90), pick(3, 300))),
            ("periodic comp", lambda X, Y: bgk.test_periodic_box_vs_oracle(orc, L.EQ_COMPRESSIBLE, X, Y), (pick(3, 70), pick(3, 300))),
            ("periodic incomp", lambda X, Y: bgk.test_periodic_box_vs_oracle(orc, L.EQ_INCOMPRESSIBLE, X, Y), (pick(3, 70), pick(3, 300))),
            ("staircase", lambda X, Y, P: bgk.test_staircase_bounce_back_vs_oracle(orc, X, Y, P), (pick(40, 110), pick(40, 140), pick(1, 3))),
            ("mrtcg rt", lambda R, C: tp.test_mrtcg_rayleigh_taylor_vs_oracle(orc, R, C), (pick(8, 80), pick(8, 280))),
            ("mrtcg ragged", lambda R, C: tp.test_mrtcg_ragged_and_tiny_grids(orc, R, C), (pick(5, 140), pick(5, 140))),
            ("rk ragged", lambda n: tp.test_rk_ragged_and_tiny_grids(orc, n), (pick(5, 270),)),
            ("kbc ragged", lambda R, C: kbc.test_tiny_and_ragged_grids(orc, R, C), (pick(3, 140), pick(3, 140))),
            ("kbc shear", lambda R, C: kbc.test_double_shear_flow_vs_oracle(orc, R, C), (pick(8, 90), pick(8, 150))),
            ("csf", lambda R, C: csf.test_csf_vs_oracle(orc, R, C), (pick(12, 100), pick(9, 270))),
            ("csf single pass", lambda R, C, rpb, kernel: csf.test_csf_single_pass_equals_three_pass_without_contraction(R, C, rpb, kernel),
             (pick(8, 150), pick(12, 400), int(rng.choice([0, 16, 32, 128])), str(rng.choice(csf.KERNELS)))),
        ]
        for name, fn, args in cases:
            try:
                fn(*args)
                print(f"round {i} {name:16s} {args} ok", flush=True)
            except Exception:  # noqa: BLE001
                failed.append((name, args))
                print(f"round {i} {name:16s} {args} FAILED\n{traceback.format_exc()}", flush=True)
    print("failed:", failed)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 3))
