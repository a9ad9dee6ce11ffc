#!/usr/bin/env python
"""One-screen summary of an `ncu --set full` report (run here, no GPU needed):  tools/ncu_summary.py report.ncu-rep [out.txt]
Duration, DRAM bytes and throughput, pipe utilisation, occupancy and its limiters, the top warp-stall reasons."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]


def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = []
    for v in vals:
        lines.append(f"== {v[col['Kernel Name']]}  grid {v[col.get('Grid Size', 0)]} block {v[col.get('Block Size', 0)]}")
        for k in KEYS:
            if k in col:
                lines.append(f"  {k:75s} {v[col[k]]:>18s} {units[col[k]]}")
        stalls = [(h, float(v[i] or 0)) for h, i in col.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        lines.append("  top warp stalls (warps stalled per issue-active cycle):")
        for h, x in sorted(stalls, key=lambda a: -a[1])[:9]:
            lines.append(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {x:8.3f}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:3])
