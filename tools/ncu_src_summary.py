#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report taken with --set full --import-source on:
    python tools/ncu_src_summary.py gpurun_out/X.ncu-rep [frames] [top]
prints, for the CUDA source lines that matter, instructions executed, shared-memory wavefronts
(total / excessive = bank conflicts) and stall samples -- per frame when `frames` is given."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines = []
    fname, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        col = {n: i for i, n in enumerate(hdr)}   # ("Source" appears twice: the last one wins, unused)

        def num(name):
            try:
                return float(r[col[name]].replace(",", ""))
            except Exception:
                return 0.0
        lines.append(dict(file=fname, line=int(r[0]), src=r[1].strip(), inst=num("Instructions Executed"),
                          wf=num("L1 Wavefronts Shared"), wfx=num("L1 Wavefronts Shared Excessive"),
                          stall=num("Warp Stall Sampling (All Samples)"), mio=num("stall_mio"), sb=num("stall_short_sb"),
                          lsb=num("stall_long_sb"), bar=num("stall_barrier"), wait=num("stall_wait"), math=num("stall_math")))
    tot = {k: sum(l[k] for l in lines) for k in ("inst", "wf", "wfx", "stall")}
    print(f"total: inst {tot['inst'] / frames:.1f}  smem wavefronts {tot['wf'] / frames:.1f} (excessive {tot['wfx'] / frames:.1f})  "
          f"stall samples {tot['stall']:.0f}   [per frame, frames={frames:g}]")
    for key, title in (("wf", "shared-memory wavefronts"), ("inst", "instructions executed"), ("stall", "stall samples")):
        print(f"--- top {top} lines by {title}")
        for l in sorted(lines, key=lambda l: -l[key])[:top]:
            if l[key] <= 0:
                break
            print(f"{l['file']}:{l['line']:<5d} inst {l['inst'] / frames:8.1f} wf {l['wf'] / frames:7.1f} x {l['wfx'] / frames:6.1f} "
                  f"stall {100 * l['stall'] / max(tot['stall'], 1):5.1f}% (mio {l['mio']:.0f} ssb {l['sb']:.0f} lsb {l['lsb']:.0f} bar {l['bar']:.0f} wait {l['wait']:.0f} math {l['math']:.0f})  | {l['src'][:90]}")


if __name__ == "__main__":
    main()
