#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into the markdown tables kept
under profiles/ (per-kernel totals and shares, then one steady-state step in launch order)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, title, steps_hint=None):
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("lbm::", "")
        rows.append((name, us, r["Grid Size"], r["Block Size"]))
    tot = OrderedDict()
    for n, us, *_ in rows:
        a = tot.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in tot.values())
    print(f"# ncu launch list — {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none --csv` on one B200; per-launch times are cold-cache and "
          "serialised, compare shares.\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for n, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {us / 1e3:.3f} | {100 * us / total:.1f}% |")
    # one steady-state step: the launches between the last two launches of the dominant kernel
    dom = max(tot.items(), key=lambda kv: kv[1][1])[0]
    idx = [i for i, r in enumerate(rows) if r[0] == dom]
    if len(idx) >= 2:
        a, b = idx[-2] + 1, idx[-1] + 1
        print("\nOne steady-state step (launch order, ending with the dominant kernel):\n")
        print("| # | kernel | grid | block | us |\n|---|---|---|---|---:|")
        for k, (n, us, g, bl) in enumerate(rows[a:b]):
            print(f"| {k} | `{n}` | {g} | {bl} | {us:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
