"""One small pair through every entry family; run under `compute-sanitizer --tool memcheck` on the GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth  # noqa: E402

import stereo_matching_cuda_b200 as S  # noqa: E402
from stereo_matching_cuda_b200 import api  # noqa: E402

L, R = synth.make_pair(300, 70, 9, channels=3, seed=1)
with S.Context(0) as ctx:
    p = api.default_params(dmin=-8, dmax=0)
    out = ctx.pipeline(L, R, p)
    gl, gr = out["gray_left"], out["gray_right"]
    cost = ctx.compute_cost(gl, gr, -8, p)
    best = np.full(gl.shape, 3e38, np.float32)
    dmap = np.zeros(gl.shape, np.float32)
    ctx.compute_guided_filter(gl, cost[:3], best, dmap, -8, api.default_params(dmin=-8, dmax=0, box_mode=S.BOX_SAT))
    occ = out["disp_left"].copy()
    ctx.detect_occlusion(occ, out["disp_right"], -108, p)
    ctx.fill_occlusion(occ, -8)
    assert np.array_equal(occ, out["filled"])
print("sanitize_small ok")
