#!/bin/bash
# A/B several builds of libstereo_b200.so on ONE box: put lib_<name>.so under gpurun_ab/ (git-ignored, travels
# with gpurun) and run `bash tools/ab.sh head variant ...`; prints ms/step, fused-kernel ms, end-to-end ms.
for round in 1 2; do
for v in "$@"; do
  SB200_LIB=$PWD/gpurun_ab/lib_$v.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['e2e']['ms_per_step'],3))"
done; done
