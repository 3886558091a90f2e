"""Developer check of the tensor-core RGB-guide kernel (k_fused_mma_rgb, SB200_RGB_KERNEL=4) against the shuffle kernel
(k_fused_cvf_rgb3, =3) and the exact-mode oracle on a few shapes, then a timing of both at 1920x1080, D=256:
    timeout 300 python tools/dev_rgb.py [quick|time]
Each kernel runs in its own context (SB200_RGB_KERNEL is read by sb200_ctx_create)."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _oracle as O  # noqa: E402
import synth  # noqa: E402
import stereo_matching_cuda_b200 as S  # noqa: E402
from stereo_matching_cuda_b200 import api  # noqa: E402


def ctx_with(kernel):
    os.environ["SB200_RGB_KERNEL"] = str(kernel)
    return S.Context(0)


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
    mode = sys.argv[1] if len(sys.argv) > 1 else ""
    orc = O.load_oracle()
    c4, c3 = ctx_with(4), ctx_with(3)
    assert c4.rgb_kernel == 4 and c3.rgb_kernel == 3
    want = ("disp_left", "disp_right", "best_left", "best_right", "filled")
    shapes = [(150, 70, -11, 0), (470, 130, -20, 0), (216, 40, -3, 0), (33, 25, -2, 0), (217, 19, 0, 0), (300, 30, 2, 9),
              (640, 256, -63, 0)]
    if mode == "time":
        shapes = shapes[:1]
    for (w, h, dmin, dmax) in shapes:
        size_d = dmax - dmin + 1
        L, R = synth.make_pair(w, h, max(size_d, 2), channels=3, seed=w + h)
        p = api.default_params(dmin=dmin, dmax=dmax, guide_mode=S.GUIDE_RGB)
        print(f"shape {w}x{h} d[{dmin},{dmax}] ...", flush=True)
        om = c4.pipeline(L, R, p, want=want)
        osf = c3.pipeline(L, R, p, want=want)
        compare("  mma vs shfl  ", om, osf)
        gl, gr = orc.rgb_to_gray(L), orc.rgb_to_gray(R)
        po = orc.params(box_mode=O.BOX_EXACT, nthreads=orc.max_threads())
        bl, dl, sl = orc.view_disparity_rgb(L, gl, gr, size_d, dmin, po, want_second=True)
        br, dr, sr = orc.view_disparity_rgb(R, gr, gl, size_d, -dmax, po, want_second=True)
        rr = {"disp_left": dl, "disp_right": dr, "best_left": bl, "best_right": br}
        compare("  mma vs oracle", om, rr)
        compare("  shfl vs oracle", osf, rr)
        for lab, best, second, o_lab in ((dl, bl, sl, "disp_left"), (dr, br, sr, "disp_right")):
            decisive = (second - best) > 2e-4
            print(f"    {o_lab}: decisive pixels that differ (mma): {np.sum((om[o_lab] != lab) & decisive)}", flush=True)
    if mode == "quick":
        return
    w, h, size_d = 1920, 1080, 256
    L, R = synth.make_pair(w, h, size_d, channels=3, seed=3)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
    res = {}
    for name, ctx in (("mma", c4), ("shfl", c3)):
        ctx.enable_timing()
        tt = []
        for _ in range(5):
            res[name] = ctx.pipeline(L, R, p, want=("disp_left", "disp_right", "best_left", "best_right"))
            tt.append(ctx.last_timing())
        print("kernel", name, {k: round(float(np.median([t[k] for t in tt[1:]])), 4) for k in tt[0]}, flush=True)
    compare("1080p mma vs shfl", res["mma"], res["shfl"])


if __name__ == "__main__":
    main()
