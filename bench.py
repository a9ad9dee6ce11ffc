#!/usr/bin/env python
"""bench.py — MLUPS of the fused D2Q9 collide+stream step on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU path, rank 0 only)

A "step" is one lattice-Boltzmann time step of the whole grid.  Workload at every N: configs[1] of
BASELINE.json — flow past a cylinder, D2Q9 BGK, compressible equilibrium, immersed-boundary cylinder
(multi-direct forcing), anti-bounce-back inlet/outlet rows, specular side columns
(test/cylinder_test.cpp of the reference), 8192 x 8192 nodes PER GPU (weak scaling: the global
grid is (8192 N) x 8192, slab-decomposed along axis 0 like test/decompose_domain.cpp).
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"))

BYTES_PER_NODE = {"bgk": 144.0}  # 9 populations x 8 B x (1 read + 1 write)
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--X", type=int, default=8192, help="rows per GPU")
    ap.add_argument("--Y", type=int, default=8192, help="columns")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="edge of the square grid the CPU baseline is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# workload description (shared by both arms)
# ---------------------------------------------------------------------------------------------
def lattice_parameters():
    """omega and u_lb exactly as params::lattice derives them from configs/parameters.toml"""
    import lbm_b200 as L

    p = L.params_from_toml(os.path.join(ROOT, "configs", "parameters.toml"), False)
    return p.omega, p.u


def cylinder_markers(X, Y):
    """Cylinder of diameter ~X/9 (the reference's X = 9 l proportion, src/params.cpp:64) centred at
    (X/4, Y/2), markers about one lattice unit apart on the circle."""
    D = max(8.0, X / 9.0)
    n = max(16, int(round(np.pi * D)))
    th = 2.0 * np.pi * np.arange(n) / n
    return X / 4.0 + 0.5 * D * np.cos(th) + 0.37, Y / 2.0 + 0.5 * D * np.sin(th) + 0.21


def write_markers_toml(path, xs, ys):
    with open(path, "w") as fh:
        fh.write("[cylinder-a]\n")
        fh.write("x = [" + ", ".join(repr(float(v)) for v in xs) + "]\n")
        fh.write("y = [" + ", ".join(repr(float(v)) for v in ys) + "]\n")


# ---------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------
def cpu_reference_mlups(edge, warmup, steps):
    """The reference's own CPU implementation of the cylinder loop (oracle/_ref, CPU libtorch, all host
    threads); falls back to the plain-C oracle port when the reference was not compiled."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    omega, u_lb = lattice_parameters()
    xs, ys = cylinder_markers(edge, edge)
    if oracle_lib.have_ref():
        ref = oracle_lib.Ref()
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "boundary.toml")
            write_markers_toml(path, xs, ys)
            sec, _ = ref.cylinder_loop(edge, edge, omega, u_lb, path, warmup, steps)
        return edge * edge / sec / 1e6, sec, "reference", ref.num_threads()
    orc = oracle_lib.Oracle()
    ib = orc.ibm_create(xs, ys)
    u = np.zeros((edge, edge, 2)); u[..., 0] = u_lb
    rho = np.ones((edge, edge, 1))
    f = orc.incomp_equilibrium(u, rho)
    for _ in range(warmup):
        orc.cylinder_step(f, u, rho, omega, u_lb, ib)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.cylinder_step(f, u, rho, omega, u_lb, ib)
    sec = (time.perf_counter() - t0) / steps
    orc.ibm_destroy(ib)
    return edge * edge / sec / 1e6, sec, "port", os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    edge = args.cpu_sample
    mlups, sec, kind, cores = cpu_reference_mlups(edge, args.warmup, args.steps)
    sample = f"{edge}x{edge} crop of the cylinder workload, {args.steps} steps after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {
        "workload": f"cylinder flow, D2Q9 BGK + IBM cylinder, ABB inlet/outlet, specular walls, {args.X}x{args.Y} nodes per GPU",
        "grid_per_gpu": [args.X, args.Y], "global_grid": [args.X * n, args.Y], "decomposition": f"{n} slab(s) along axis 0",
        "l2": "inputs larger than L2 (2 x 4.8 GB of populations per GPU), no flush needed",
        "reference_driver": "test/cylinder_test.cpp",
    }


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of the interior kernel from the committed ncu capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            return json.load(fh).get("k_bgk_interior_bytes_per_launch_8192x8192")
    except Exception:
        return None


def run_b200_arm(args):
    import torch

    import lbm_b200 as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    omega, u_lb = lattice_parameters()
    Xg, Y = args.X * world, args.Y
    x0, x1 = rank * args.X, (rank + 1) * args.X
    cfg = L.default_config(model=L.MODEL_BGK, X=Xg, Y=Y, x0=x0, x1=x1, device=local, omega=omega,
                           equilibrium=L.EQ_COMPRESSIBLE, force=L.FORCE_IBM)
    d = L.Domain(cfg)
    if world > 1:
        ident = [L.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        d.comm_init(ident[0], world, rank)
    d.preset_free_stream(u_lb, 0.0)
    xs, ys = cylinder_markers(args.X, Y)  # the body sits in rank 0's slab
    if rank == 0:
        d.ibm_set_markers(xs, ys)

    # initial state from pinned host memory (the drivers' incomp_equilibrium(u=(u_lb,0), rho=1))
    N = args.X * Y
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    cx = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1], dtype=np.float64)
    f_host_t = torch.empty((args.X, Y, 9), dtype=torch.float64, pin_memory=True)
    f_host = f_host_t.numpy()
    f_host[...] = (1.0 + 3.0 * cx * u_lb) * w
    rho_host_t = torch.empty((args.X, Y, 1), dtype=torch.float64, pin_memory=True)
    u_host_t = torch.empty((args.X, Y, 2), dtype=torch.float64, pin_memory=True)

    def barrier():
        d.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d.set_f(f_host)
    d.step(args.warmup)
    barrier()

    # ---- device-resident timed region: exactly K steps, CUDA events on the domain's stream
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = d.kernel_launches()
    d.profile_enable(True)
    barrier()
    d.step(args.steps)
    d.synchronize()
    ms = d.last_step_ms()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = d.kernel_launches() - launches0
    prof = {name: d.profile_read(cls) for name, cls in
            [("interior", L.PROF_INTERIOR), ("boundary", L.PROF_BOUNDARY), ("fixup", L.PROF_FIXUP),
             ("ghost", L.PROF_GHOST), ("ibm", L.PROF_IBM)]}
    d.profile_enable(False)
    ms = max_over_ranks(ms)
    mlups = (Xg * Y) * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers: import f (H2D), K steps, export rho,u (D2H)
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        d.set_f(f_host)
        d.step(args.steps)
        lib = L.load()
        dp = ctypes.POINTER(ctypes.c_double)
        rc = lib.lbm_get_moments(d.h, 0, rho_host_t.numpy().ctypes.data_as(dp), u_host_t.numpy().ctypes.data_as(dp))
        assert rc == 0, lib.lbm_last_error()
        d.synchronize()
        sec = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": (Xg * Y) * args.steps / sec / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": N * 9 * 8 / args.steps, "d2h_bytes_per_step": N * 3 * 8 / args.steps,
               "what": f"lbm_set_f from pinned host + lbm_step({args.steps}) + lbm_get_moments to pinned host, per rank",
               "seconds": sec, "rho_mean": float(rho_host_t.mean())}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_bgk_interior: one launch over the early rows + one over the
    # bulk rows per step; both timed with CUDA events on the main stream inside the timed region above)
    peak, peak_src = measured_hbm_peak()
    int_ms, int_n = prof["interior"]
    nodes_per_step = args.X * (2 * ((Y - 3) // 2))   # nodes the interior kernel owns (edge columns are listed nodes)
    launches_per_step = int_n / args.steps if args.steps else 0
    achieved = BYTES_PER_NODE["bgk"] * nodes_per_step * args.steps / (int_ms * 1e-3) / 1e9 if int_ms > 0 else None
    traffic = ncu_traffic_per_launch()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": traffic,
                "kernel": "k_bgk_interior<PULL,COMP,IBM>", "bytes_per_node": BYTES_PER_NODE["bgk"],
                "algorithmic_bytes_per_step": BYTES_PER_NODE["bgk"] * nodes_per_step,
                "launches_per_step": launches_per_step, "kernel_ms_per_step": int_ms / args.steps,
                "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0 if achieved else None,
                "whole_step_frac": BYTES_PER_NODE["bgk"] * mlups * 1e6 / 1e9 / peak,
                "whole_step_frac_of_nominal_8TBs": BYTES_PER_NODE["bgk"] * mlups * 1e6 / 1e9 / 8000.0,
                "share_of_step": int_ms / ms,
                "side_stream_spans_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if k != "interior"},
                "note": "achieved = 144 B x interior nodes per step / summed duration of the step's interior launches; "
                        "listed nodes, stages, ghost rows and the IBM pre-pass run on a side stream under the bulk launch"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        c_mlups, c_sec, kind, cores = cpu_reference_mlups(args.cpu_sample, 1, 5)
        cpu_baseline = {"value": c_mlups, "unit": "MLUPS", "cores": cores, "kind": kind,
                        "sample": f"{args.cpu_sample}x{args.cpu_sample} crop of the cylinder workload, 5 steps after 1 warm-up",
                        "ms_per_step": c_sec * 1e3}

    line = {
        "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
