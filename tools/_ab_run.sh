cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=stereo_matching_cuda_b200/libstereo_b200.so
AB_GUIDE=gray python tools/ab_rgb.py prev=gpurun_ab/lib_prev.so new=$L
python tools/ab_rgb.py prev=gpurun_ab/lib_prev.so new=$L
for l in gpurun_ab/lib_prev.so $L; do SB200_LIB=$PWD/$l python bench.py --no-legs --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l', d['ms_per_step'], d['roofline']['other_kernels_ms'], d['gpu_launches'])"; done
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shapes or tiny or strip or tsukuba" 2>&1 | tail -2
