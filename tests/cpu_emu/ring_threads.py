"""TEST INFRASTRUCTURE (see cuda_emu.hpp): tests/mp_nccl_check.py's checks of the slab ring with one THREAD per rank, on
the emulated device, over the in-process NCCL stand-in (fake_nccl.cpp; LBM_NCCL_LIB).  What the GPU box runs under
torchrun with real NCCL — ghost rows, moment / normal halos, pressure packets across the ring, an immersed body whose ROI
crosses the cuts — runs here at any ring size; a rank that waits for a message nobody sends aborts the run instead of
hanging it.          usage: ring_threads.py WORLD"""
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"), os.path.join(ROOT, "tests")]
BUILD = os.path.join(HERE, "_build_asan" if os.environ.get("LBM_EMU_ASAN") == "1" else "_build")  # SAN=1 build: run with libasan preloaded
os.environ["LBM_NCCL_LIB"] = os.path.join(BUILD, "libnccl_emu.so")

import numpy as np  # noqa: E402

import lbm_b200 as L  # noqa: E402

L.LIB_PATH = os.path.join(BUILD, "liblbm_b200_emu.so")
import mp_nccl_check  # noqa: E402


def main(world):
    L.load()
    sync = threading.Barrier(world)
    shared = {"id": None, "parts": [None] * world}
    results = [None] * world

    def worker(rank):
        def fresh_id():
            if rank == 0:
                shared["id"] = L.comm_unique_id()
            sync.wait()
            ident = shared["id"]
            sync.wait()
            return ident

        def gather(a):
            shared["parts"][rank] = a
            sync.wait()
            out = np.concatenate(shared["parts"], axis=0)
            sync.wait()
            return out

        try:
            results[rank] = mp_nccl_check.run_checks(rank, world, 0, fresh_id, gather, sync.wait, extra=True)
        except BaseException as e:  # noqa: BLE001  (a failed rank must not leave the others at a barrier)
            results[rank] = [f"rank {rank}: {type(e).__name__}: {e}"]
            sync.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    failures = [f for r in results if r for f in r]
    print("failures:", failures)
    return 1 if failures or any(r is None for r in results) else 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 2))
