"""Timeline of role B of block 0 of k_fused_mma (library built with -DMMA_TIMELINE: tools/build_variant.sh tlg fused_mma.cu
-DMMA_TIMELINE): per iteration, cycles in barrier waits / Tensor-Memory loads / arithmetic + stores + fence.
    SB200_LIB=gpurun_ab/lib_tlg.so python tools/timeline_gray.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
import stereo_matching_cuda_b200 as S  # noqa: E402
from stereo_matching_cuda_b200 import api  # noqa: E402

w, h, size_d = 1920, 1080, 256
L, R = synth.make_pair(w, h, size_d, seed=0)
p = api.default_params(dmin=-(size_d - 1), dmax=0)
with S.Context(0) as ctx:
    for _ in range(2):
        ctx.pipeline(L, R, p, want=("disp_left",))
    N = 4096
    buf = np.zeros((6, N, 4), np.int64)
    lib = ctx.lib
    lib.sb200_debug_timeline_gray.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert lib.sb200_debug_timeline_gray(buf.ctypes.data, buf.nbytes) == 0
t = buf[0, :, 0]
n = int((t > 0).sum())
dt = np.diff(t[:n])
print(f"role B: iterations {n}, period mean {dt.mean():.0f} med {np.median(dt):.0f} p90 {np.percentile(dt, 90):.0f}")
print(f"  per iteration: barrier waits {buf[0, :n, 1].mean():.0f}, TMEM loads (2 halves) {buf[0, :n, 2].mean():.0f}, "
      f"arithmetic + ring/B2 stores + fence + arrive {buf[0, :n, 3].mean():.0f}")
