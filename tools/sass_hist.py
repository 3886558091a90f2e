#!/usr/bin/env python
"""Opcode histogram of every kernel of the product library, the evidence that the hot kernels are hand-written sm_100a
code (tensor-core MMAs with TMEM operands, TMEM loads/stores, TMA bulk copies, mbarriers, setmaxnreg, packed-half math):
    python tools/sass_hist.py [stereo_matching_cuda_b200/libstereo_b200.so] > profiles/r2_sass_histogram.txt"""
import re
import subprocess
import sys
from collections import Counter

lib = sys.argv[1] if len(sys.argv) > 1 else "stereo_matching_cuda_b200/libstereo_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "SYNCS", "USETMAXREG", "ELECT", "FHFMA", "FHADD", "HFMA2",
       "HADD2", "HMUL2", "HMNMX2", "F2FP", "FFMA", "FADD", "FMUL", "FSEL", "FSETP", "SHFL", "LDS", "STS", "LDG", "STG", "MEMBAR", "FENCE")
cur, hist = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
    if cur and m:
        op = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip()).split()[0].split(".")[0]
        hist[cur][op] += 1
print(f"SASS opcode counts per kernel of {lib} (cuobjdump -sass; static instruction counts)\n")
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    c = hist[k]
    short = re.sub(r"^_ZN\d+_GLOBAL__N__\w+?\d+", "", k)
    name = re.search(r"(k_[a-z0-9_]+)", k)
    print(f"{name.group(1) if name else short}: {sum(c.values())} instructions")
    print("    " + "  ".join(f"{op}:{c[op]}" for op in KEY if c[op]))
