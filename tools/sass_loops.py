#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in a .o/.so with instruction counts per loop body:
    python tools/sass_loops.py stereo_matching_cuda_b200/csrc/fused_cvf_rgb3.o k_fused_cvf_rgb3
Used to balance the warp roles of the fused kernels before spending GPU time."""
import re
import subprocess
import sys
from collections import Counter

obj, kern = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, ins = None, []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and f"{len(kern)}{kern}" in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f"{kern}: {len(ins)} instructions")
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            body = ins[addr_index[tgt]:i + 1]
            if len(body) < minlen:
                continue
            ops = Counter()
            for _, tt in body:
                op = re.sub(r"^@!?U?P\d+\s+", "", tt).split()[0].split(".")[0]
                ops[op] += 1
            top = " ".join(f"{k}:{v}" for k, v in ops.most_common(14))
            print(f"  loop [{tgt:#x}, {a:#x}] {len(body)} instr  {top}")
