#!/bin/bash
# one ncu --set full capture of the RGB fused kernel at 1080p D=256 -> gpurun_out/prof_$1.ncu-rep
set -u
name=${1:-rgb}
lib=${2:-stereo_matching_cuda_b200/libstereo_b200.so}
timeout 100 python tools/ab_rgb.py head=$lib
SB200_LIB=$(realpath $lib) timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_fused_mma_rgb -s 1 -c 1 -f \
  -o gpurun_out/prof_$name python tools/ab_rgb.py --child /tmp/ab_rgb_pair.npz > gpurun_out/ncu_$name.log 2>&1
tail -3 gpurun_out/ncu_$name.log
