#!/usr/bin/env python
"""gpurun_out/<tag>_* (what tools/profile_round.sh leaves) -> profiles/<tag>_*:
  <tag>_full_X.raw.csv  (ncu --page raw --csv: one wide row per kernel)  -> <tag>_full_X.csv  (metric, unit, value)
  <tag>_launches_bench.csv (ncu --metrics gpu__time_duration.sum --csv)  -> id, kernel (shortened), block, grid, ns
  bench lines / probe outputs are copied as they are.
usage: python tools/summarize_profiles.py r1e"""
import csv
import glob
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def long_form(src, dst):
    rows = [r for r in csv.reader(open(src)) if r]
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for n, u, v in zip(names, units, vals):
            w.writerow([n, u, v])


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r]
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names = rows[hdr]
    col = {n: names.index(n) for n in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Value")}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "gpu__time_duration.sum [ns]"])
        for r in rows[hdr + 2:]:
            if len(r) <= col["Metric Value"]:
                continue
            k = r[col["Kernel Name"]]
            k = re.sub(r"\(.*$", "", k)                      # drop the argument list
            k = re.sub(r"<.*$", "", k) if k.startswith("void at::") else k
            w.writerow([r[col["ID"]], k, r[col["Block Size"]], r[col["Grid Size"]], r[col["Metric Value"]].replace(",", "")])


def main():
    tag = sys.argv[1]
    for p in sorted(glob.glob(os.path.join(SRC, tag + "_*"))):
        b = os.path.basename(p)
        if b.endswith(".raw.csv"):
            long_form(p, os.path.join(DST, b.replace(".raw.csv", ".csv")))
        elif b.endswith("_launches_bench.csv"):
            launches(p, os.path.join(DST, b))
        elif b.endswith(".json") or b.endswith(".jsonl") or (b.endswith(".txt") and "_src_" in b):
            shutil.copy(p, os.path.join(DST, b))
        else:
            continue
        print("profiles/" + b.replace(".raw.csv", ".csv"))


if __name__ == "__main__":
    main()
