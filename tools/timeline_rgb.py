"""Timeline of block 0 of k_fused_mma_rgb (library built with -DRGBM_TIMELINE: tools/build_variant.sh tl fused_mma_rgb.cu
-DRGBM_TIMELINE): per role and iteration, the start clock and the cycles spent waiting on each kind of barrier.
    SB200_LIB=gpurun_ab/lib_tl.so python tools/timeline_rgb.py"""
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
L, R = synth.make_pair(w, h, size_d, channels=3, seed=3)
p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
with S.Context(0) as ctx:
    for _ in range(2):
        ctx.pipeline(L, R, p, want=("disp_left",))
    N = 4096
    buf = np.zeros((6, N, 4), np.int64)
    lib = ctx.lib
    lib.sb200_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = lib.sb200_debug_timeline(buf.ctypes.data, buf.nbytes)
    assert rc == 0, rc
names = {0: ("B", "s_full", "d1_full", "b2_empty"), 1: ("C", "gc_full", "d2_full", "-"), 2: ("A", "a_full+b1_empty", "loads+math+ring", "B1 stores+fence"),
         3: ("MMA1", "b1_full", "d1_empty", "-"), 4: ("MMA2", "b2_full", "d2_empty", "-")}
np.save("gpurun_out/timeline_rgb.npy", buf)
t00 = buf[2, 0, 0]
for role, nm in names.items():
    t = buf[role, :, 0]
    ok = t > 0
    n = int(ok.sum())
    dt = np.diff(t[:n])
    print(f"{nm[0]:5s} iterations {n}  period mean {dt.mean():.0f} med {np.median(dt):.0f} p90 {np.percentile(dt, 90):.0f} max {dt.max()}"
          f"  waits/iter: {nm[1]} {buf[role, :n, 1].mean():.0f}  {nm[2]} {buf[role, :n, 2].mean():.0f}  {nm[3]} {buf[role, :n, 3].mean():.0f}")
# lag between the roles at the same iteration index
n = int((buf[:5, :, 0] > 0).all(axis=0).sum())
for a, b in ((2, 3), (3, 0), (0, 4), (4, 1)):
    lag = buf[b, :n, 0] - buf[a, :n, 0]
    print(f"lag {names[a][0]}->{names[b][0]}: mean {lag.mean():.0f} med {np.median(lag):.0f}")
# print a window of iterations
for K in list(range(300, 312)):
    print(K, " ".join(f"{names[r][0]}:{buf[r, K, 0] - t00:8d}+{buf[r, K, 1]:5d}/{buf[r, K, 2]:5d}/{buf[r, K, 3]:5d}" for r in range(5)))
