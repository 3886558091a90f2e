#!/bin/bash
# Build gpurun_ab/lib_<name>.so = the in-tree objects with ONE source recompiled with extra flags:
#   bash tools/build_variant.sh hs1 fused_cvf_rgb3.cu -DRGB3_HS1=1
# (run `python -m stereo_matching_cuda_b200.build` first so the other objects exist)
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/../stereo_matching_cuda_b200/csrc"
mkdir -p ../../gpurun_ab
objs=""
for f in api stage_kernels fused_cvf fused_mma fused_mma_rgb fused_cvf_rgb fused_cvf_rgb3; do
  if [ "$f.cu" == "$src" ]; then
    nvcc -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC \
      --expt-relaxed-constexpr "$@" -c $src -o /tmp/variant_$name.o 2>&1 | grep -v deprecated || true
    objs="$objs /tmp/variant_$name.o"
  else
    objs="$objs $f.o"
  fi
done
nvcc -ccbin /usr/bin/g++ -shared -o ../../gpurun_ab/lib_$name.so $objs 2>&1 | grep -v deprecated || true
ls -la ../../gpurun_ab/lib_$name.so
