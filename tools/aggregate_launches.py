#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for ONE training step.

The last complete step is delimited by the optimizer kernel (`adamw_kernel` / `sgd_nesterov_kernel`, one
launch per step): everything after the second-to-last optimizer launch up to and including the last one.
usage: aggregate_launches.py launches.csv > one_step.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = val / 1e6 if unit in ("ns", "nsecond") else (val / 1e3 if unit in ("us", "usecond") else val)
        rows.append((r["Kernel Name"], ms))
    opt = [i for i, (k, _) in enumerate(rows) if "adamw_kernel" in k or "sgd_nesterov_kernel" in k]
    if len(opt) < 2:
        sys.exit("need at least two optimizer launches to delimit a step")
    step = rows[opt[-2] + 1:opt[-1] + 1]
    agg = OrderedDict()
    for k, ms in step:
        k = re.sub(r"\(.*$", "", k)
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ms)
    total = sum(t for _, t in agg.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches_per_step", "total_ms", "share"])
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, f"{t:.3f}", f"{t / total:.3f}"])
    w.writerow(["TOTAL", len(step), f"{total:.3f}", "1.000"])


if __name__ == "__main__":
    main(sys.argv[1])
