#!/bin/bash
# Round-end measurement on one B200 box: ncu captures of the two fused kernels -> summaries + the JSON bench.py quotes, the
# FMA-chain micro-benchmark, the launch list of the bench command, a compute-sanitizer attempt and the bench lines.
# Everything lands in gpurun_out/final/ (copy what should be judged into profiles/).
set -u
out=gpurun_out/final; mkdir -p $out
lib=stereo_matching_cuda_b200/libstereo_b200.so
./tools/fma_chain > $out/fma_chain.txt 2>&1
AB_GUIDE=gray timeout 100 python tools/ab_rgb.py head=$lib > $out/ab_gray.log 2>&1
AB_GUIDE=gray timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_fused_mma -s 1 -c 1 -f -o $out/prof_gray \
  python tools/ab_rgb.py --child /tmp/ab_rgb_pair.npz > $out/ncu_gray.log 2>&1
timeout 100 python tools/ab_rgb.py head=$lib > $out/ab_rgb.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_fused_mma_rgb -s 1 -c 1 -f -o $out/prof_rgb \
  python tools/ab_rgb.py --child /tmp/ab_rgb_pair.npz > $out/ncu_rgb.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-legs > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $out/launches_rgb.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-legs --guide rgb > /dev/null 2>&1
(timeout 150 compute-sanitizer --tool memcheck python tests/sanitize_small.py 2>&1 | tail -15) > $out/sanitizer.txt
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/fma_chain.txt $out/ab_gray.log $out/ab_rgb.log; tail -3 $out/sanitizer.txt; head -c 600 $out/bench_n1.json
