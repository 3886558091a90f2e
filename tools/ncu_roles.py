#!/usr/bin/env python
"""Stall breakdown per warp role of a fused-MMA kernel capture: python tools/ncu_roles.py rep lo:hi:name ...
(address ranges relative to the kernel start, hex).  Barrier spin sites (BRA dominated by long_sb) are counted apart."""
import collections
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    stall = [(k, i) for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
    res = []
    for r in rows[2:]:
        res.append((int(r[0], 16), r[1].strip(), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]),
                    {k[6:]: int(r[i]) for k, i in stall if r[i].isdigit() and int(r[i]) > 0}))
    base = res[0][0]
    return [(a - base, s, smp, e, st) for a, s, smp, e, st in res]


def main():
    res = load(sys.argv[1])
    for spec in sys.argv[2:]:
        lo, hi, name = spec.split(":")
        lo, hi = int(lo, 16), int(hi, 16)
        tot = collections.Counter()
        spin = n = ex = 0
        for a, s, smp, e, st in res:
            if lo <= a < hi:
                if "BRA" in s and st.get("long_sb", 0) > 0.8 * smp and smp > 200:
                    spin += smp
                    continue
                tot.update(st)
                n += smp
                ex += e
        print("%-5s busy samples %6d  barrier-spin %6d  warp-instr %10d  %s" % (name, n, spin, ex, dict(tot.most_common(9))))


if __name__ == "__main__":
    main()
