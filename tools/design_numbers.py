#!/usr/bin/env python
"""Prints the round's results table of DESIGN.md section 6 from the committed bench records (profiles/r02_bench_*.json)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda n: json.load(open(os.path.join(ROOT, "profiles", n)))  # noqa: E731
j, n2, n8 = P("r02_bench_n1_full_final_lib.json"), P("r02_bench_n2_final_lib.json"), P("r02_bench_n8.json")  # N = 1, 2: the last library of the round
o = j["other_workloads"]
names = {"poiseuille": ("poiseuille (configs[0])", "2700x2100"), "mrtcg_rt": ("mrtcg_rt (configs[2])", "16384^2"),
         "rk_droplet": ("rk_droplet (configs[3])", "4096^2"), "sedimentation": ("sedimentation (configs[4], the reference driver)", "4096x8192"),
         "sedimentation_ibm": ("sedimentation_ibm (configs[4] as worded)", "4096x8192"), "kbc_shear": ("kbc_shear (§8c)", "8192^2"),
         "csf_rt": ("csf_rt (§8d, single pass)", "8192^2"), "cylinder_bb": ("cylinder_bb (configs[1] read literally)", "8192^2"),
         "mrtcg_rt_weak": ("mrtcg_rt_weak (one slab of the ring)", "8192x16384")}
print("| workload | grid | GLUPS | kernel frac | whole step | CUDA graph |\n|---|---|---:|---:|---:|---|")
print(f"| cylinder (headline, configs[1]) | 8192^2 | {j['value'] / 1e3:.2f} | {j['roofline']['frac']:.3f} | {j['roofline']['whole_step_frac_per_gpu']:.3f} | no |")
for k, (n, g) in names.items():
    v = o[k]
    r = v["roofline"]
    print(f"| {n} | {g} | {v['value'] / 1e3:.2f} | {r['frac']:.3f} | {r['whole_step_frac_per_gpu']:.3f} | {'yes' if v['cuda_graph'] else 'no'} |")
e = j["e2e"]
print(f"\ncpu_baseline {j['cpu_baseline']['value']:.2f} MLUPS on {j['cpu_baseline']['cores']} threads; e2e {e['value'] / 1e3:.1f} GLUPS (blocking {e['blocking']['value'] / 1e3:.1f})")
for n, r in ((2, n2), (8, n8)):
    w = r["other_workloads"]["mrtcg_rt_weak"]["value"]
    print(f"N={n}: cylinder {r['value'] / 1e3:.1f} GLUPS ({r['value'] / n / j['value'] * 100:.1f} %), mrtcg_rt_weak {w / 1e3:.1f} ({w / n / o['mrtcg_rt_weak']['value'] * 100:.1f} %), "
          f"ring_parity green={r['ring_parity']['green']} in {r['ring_parity']['seconds']:.0f} s")
