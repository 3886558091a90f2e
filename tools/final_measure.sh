#!/bin/bash
# Round-end measurement on one B200 box: ncu captures of both fused kernels -> summaries + the JSON bench.py quotes,
# then the bench lines (gray headline, RGB guide), the launch list of the bench command and the GPU test suite.
# Everything lands in gpurun_out/final/ (copy what should be judged into profiles/).
set -u
out=gpurun_out/final; mkdir -p $out
lib=stereo_matching_cuda_b200/libstereo_b200.so
AB_GUIDE=gray timeout 100 python tools/ab_rgb.py head=$lib > $out/ab_gray.log 2>&1
AB_GUIDE=gray timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused_cvf -s 1 -c 1 -f -o $out/prof_gray \
  python tools/ab_rgb.py --child /tmp/ab_rgb_pair.npz > $out/ncu_gray.log 2>&1
timeout 100 python tools/ab_rgb.py head=$lib > $out/ab_rgb.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused_cvf_rgb3 -s 1 -c 1 -f -o $out/prof_rgb3 \
  python tools/ab_rgb.py --child /tmp/ab_rgb_pair.npz > $out/ncu_rgb3.log 2>&1
python tools/ncu_summary.py $out/prof_gray.ncu-rep $out/prof_fused_summary.txt --json profiles/fused_ncu.json > /dev/null 2>&1
python tools/ncu_summary.py $out/prof_rgb3.ncu-rep $out/prof_fused_rgb3_summary.txt --json profiles/fused_rgb3_ncu.json > /dev/null 2>&1
cp profiles/fused_ncu.json profiles/fused_rgb3_ncu.json $out/
python bench.py --steps 20 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --guide rgb --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_rgb.json 2> $out/bench_rgb.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > $out/tests_gpu.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $out/smoke.log
cat $out/ab_gray.log $out/ab_rgb.log $out/tests_gpu.log $out/smoke.log; cut -c1-300 $out/bench_n1.json
