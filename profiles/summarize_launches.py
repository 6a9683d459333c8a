#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: device time per kernel over ONE step of the
genome-wide scan (the launches between two consecutive first-unit prefilter launches are one unit; a step is
`--units` units), as shares of the step's kernel time.  Per-launch times under ncu are cold-cache and serialised:
the SHARES are what is compared with bench.py's event-timed phases.

    python profiles/summarize_launches.py gpurun_out/r2_launches.csv --units 47
"""
import argparse
import collections
import csv
import re


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.split("::")[-1] if "cub" not in name else "cub::" + name.split("::")[-1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--units", type=int, default=47)
    args = ap.parse_args()
    rows = []
    with open(args.csv, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e6))
    scans = [i for i, (k, _) in enumerate(rows) if k == "prefilter_tc_kernel"]
    print(f"{len(rows)} launches, {len(scans)} prefilter_tc_kernel launches")
    if len(scans) <= args.units:
        window = rows[scans[0]:] if scans else rows
        print("fewer prefilter launches than one step: summarising everything from the first one")
    else:
        # skip the cutoff build's scan (the first prefilter launch belongs to msb_score_select when cutoffs are built here)
        first = scans[1] if len(scans) > args.units + 1 else scans[0]
        k0 = scans.index(first)
        window = rows[first:scans[k0 + args.units]] if k0 + args.units < len(scans) else rows[first:]
    per = collections.OrderedDict()
    for k, ms in window:
        a = per.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in per.values())
    print(f"one step = {len(window)} launches, {total:.2f} ms of kernel time under ncu")
    for k, (n, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:45s} {n:6d} launches {ms:10.3f} ms  {100 * ms / total:5.1f} %")


if __name__ == "__main__":
    main()
