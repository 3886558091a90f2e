#!/usr/bin/env python
"""Per-source-line view of one ncu capture (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [min_share_pct]
Prints, for every CUDA source line with at least min_share_pct of the warp-state samples, the samples, the warp-level
instructions executed and the dominant stall reasons."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = None
    head = None
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            head = r
            ix = {k: i for i, k in enumerate(head)}
            stall_cols = [(k, i) for i, k in enumerate(head) if k.startswith("stall_")]
            continue
        if head is None or not r or r[0] == "":
            continue  # SASS rows belong to the line above; the line row already aggregates them
        try:
            smp = int(r[ix["# Samples"]])
            ex = int(r[ix["Instructions Executed"]])
        except (ValueError, KeyError):
            continue
        st = sorted(((int(r[i]), k[6:]) for k, i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:4]
        lines.append((smp, ex, fname, r[0], r[1].strip()[:110], st))
    total = sum(x[0] for x in lines)
    tot_ex = sum(x[1] for x in lines)
    print("samples %d, warp instructions %d" % (total, tot_ex))
    for smp, ex, f, ln, src, st in sorted(lines, reverse=True):
        if smp < thr / 100 * total:
            break
        print("%5.1f%% %7d ex %10d  %s:%s  %s\n         %s" % (100 * smp / total, smp, ex, f, ln, src, " ".join("%s:%d" % (k, v) for v, k in st)))


if __name__ == "__main__":
    main()
