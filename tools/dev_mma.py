"""Developer check of the tensor-core fused kernel (k_fused_mma) against the warp-shuffle kernel (k_fused_cvf) and the
exact-mode oracle on a few shapes, then a timing of both at 1920x1080, D=256.  Run on the GPU box:
    timeout 300 python tools/dev_mma.py [quick]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _oracle as O  # noqa: E402
import synth  # noqa: E402
import stereo_matching_cuda_b200 as S  # noqa: E402
from stereo_matching_cuda_b200 import api  # noqa: E402


def compare(tag, a, b, keys=("disp_left", "disp_right", "best_left", "best_right")):
    msg = [tag]
    for k in keys:
        if k.startswith("disp"):
            msg.append(f"{k} agree {(a[k] == b[k]).mean():.6f}")
        else:
            err = np.abs(a[k] - b[k])
            rel = err / np.maximum(np.abs(b[k]), 1e-2)
            msg.append(f"{k} max rel(1e-2 floor) {rel.max():.2e} abs {err.max():.2e}")
    print("  ".join(msg), flush=True)


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    orc = O.load_oracle()
    ctx = S.Context(0)
    want = ("disp_left", "disp_right", "best_left", "best_right", "filled")
    shapes = [(384, 288, -15, 0), (500, 150, -33, 0), (230, 90, -20, 0), (300, 30, 2, 9), (19, 19, -2, 0), (97, 11, 0, 0),
              (640, 256, -63, 0)]
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        shapes = shapes[:1]
    for (w, h, dmin, dmax) in shapes:
        size_d = dmax - dmin + 1
        L, R = synth.make_pair(w, h, max(size_d, 2), seed=w + h)
        p = api.default_params(dmin=dmin, dmax=dmax)
        print(f"shape {w}x{h} d[{dmin},{dmax}] ...", flush=True)
        ctx.gray_kernel = 1
        t0 = time.time()
        om = ctx.pipeline(L, R, p, want=want)
        t1 = time.time()
        ctx.gray_kernel = 0
        osf = ctx.pipeline(L, R, p, want=want)
        print(f"shape {w}x{h} d[{dmin},{dmax}]  mma call {1e3 * (t1 - t0):.1f} ms", flush=True)
        compare("  mma vs shfl  ", om, osf)
        ref = orc.pipeline_gray(L, R, dmin, size_d, orc.params(box_mode=O.BOX_EXACT, nthreads=orc.max_threads()), want_second=True)
        rr = {"disp_left": ref["dL"], "disp_right": ref["dR"], "best_left": ref["bestL"], "best_right": ref["bestR"]}
        compare("  mma vs oracle", om, rr)
        compare("  shfl vs oracle", osf, rr)
        for lab, best, second, o_lab in (("dL", "bestL", "secondL", "disp_left"), ("dR", "bestR", "secondR", "disp_right")):
            decisive = (ref[second] - ref[best]) > 2e-4
            bad = np.sum((om[o_lab] != ref[lab]) & decisive)
            print(f"    {o_lab}: decisive pixels that differ (mma): {bad}", flush=True)
    if quick:
        return
    import torch

    w, h, size_d = 1920, 1080, 256
    L, R = synth.make_pair(w, h, size_d, seed=0)
    p = api.default_params(dmin=-(size_d - 1), dmax=0)
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    outs = {k: torch.empty((h, w), dtype=torch.float32, device="cuda") for k in ("disp_left", "disp_right", "best_left", "best_right", "filled")}
    ctx.set_stream(torch.cuda.current_stream())
    ctx.enable_timing(True)
    res = {}
    for which in (1, 0):
        ctx.gray_kernel = which
        for _ in range(3):
            ctx.pipeline_dev(dl, dr, 1, w, h, outs, p)
        torch.cuda.synchronize()
        tt = []
        for _ in range(5):
            ctx.pipeline_dev(dl, dr, 1, w, h, outs, p)
            torch.cuda.synchronize()
            tt.append(ctx.last_timing())
        print("kernel", "mma" if which else "shfl", {k: round(float(np.median([t[k] for t in tt])), 4) for k in tt[0]}, flush=True)
        res[which] = {k: v.cpu().numpy() for k, v in outs.items()}
    compare("1080p mma vs shfl", res[1], res[0])


if __name__ == "__main__":
    main()
