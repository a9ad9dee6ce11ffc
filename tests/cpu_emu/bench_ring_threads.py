"""TEST INFRASTRUCTURE (see cuda_emu.hpp): bench.py's `ring_parity` leg — every model family over a ring of N ranks against
the monolithic run — with one THREAD per rank on the emulated device over the in-process NCCL stand-in, so the leg the
multi-GPU bench runs before its timed region is exercised here at any ring size.     usage: bench_ring_threads.py WORLD"""
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"), os.path.join(ROOT, "tests")]
BUILD = os.path.join(HERE, "_build_asan" if os.environ.get("LBM_EMU_ASAN") == "1" else "_build")
os.environ["LBM_NCCL_LIB"] = os.path.join(BUILD, "libnccl_emu.so")

import numpy as np  # noqa: E402

import lbm_b200 as L  # noqa: E402

L.LIB_PATH = os.path.join(BUILD, "liblbm_b200_emu.so")
import bench  # noqa: E402


def main(world):
    L.load()
    sync = threading.Barrier(world)
    shared = {"id": None, "parts": [None] * world}
    results = [None] * world

    class ThreadCtx:
        def __init__(self, rank):
            self.L, self.rank, self.world, self.local, self.anchor = L, rank, world, 0, None

        def fresh_id(self):
            if self.rank == 0:
                shared["id"] = L.comm_unique_id()
            sync.wait()
            ident = shared["id"]
            sync.wait()
            return ident

        def gather_rows(self, a):
            shared["parts"][self.rank] = a
            sync.wait()
            out = np.concatenate(shared["parts"], axis=0)
            sync.wait()
            return out

        def barrier(self, d=None):
            if d is not None:
                d.synchronize()
            sync.wait()

    def worker(rank):
        try:
            results[rank] = bench.ring_parity(ThreadCtx(rank)) or {}
        except BaseException as e:  # noqa: BLE001  (a failed rank must not leave the others at a barrier)
            results[rank] = {"error": f"rank {rank}: {type(e).__name__}: {e}"}
            sync.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    print("ring_parity:", results[0])
    errors = [r["error"] for r in results if r and "error" in r]
    print("errors:", errors)
    return 0 if not errors and results[0] and results[0].get("green") else 1


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 2))
