"""Command-line driver: the reference's main() (main.cu:37-214) as one command.

    python -m stereo_matching_cuda_b200.cli LEFT.png RIGHT.png OUTDIR [--dmin -15 --dmax 0] [--fused] [--subpixel]

Reads a rectified pair, runs the pipeline on GPU 0 and writes the same 12 PNGs main.cu:162-181 writes
(image_left/right, image_mean_left/right, best_costl/r, cost_lminus15 / cost_rminus15 (slice 0 of each
volume), occlu_mapl, disparity_mapl/r, occlu_mapl_filled).  Default is the stage-by-stage path in SAT
mode, whose float32 arithmetic is the reference's operation by operation, so on the Tsukuba pair the
output files decode to exactly the reference's checked-in PNGs; --fused uses the fused kernel instead.
--subpixel (beyond the reference, implies the fused kernel) adds disparity_mapl_subpixel.png, the sub-pixel refined and
filled left disparity through the same write_mat normalisation, and disparity_mapl_subpixel.npy with the float values.
"""
import argparse
import os

import numpy as np

from . import BOX_SAT, Context
from .api import default_params


def write_mat(mat):
    """main.cu:13-35 on the host (numpy): min/max scan with its `else if (<=)` quirk, then (v-min)*255/(max-min)
    truncated.  The driver below uses the device version (Context.write_mat); this one is kept as its cross-check."""
    flat = np.ascontiguousarray(mat, np.float32).ravel()
    mx, mn = np.float32(-150000000.0), np.float32(150000000.0)
    # sequential semantics: a value that raises the max is never considered for the min
    run_max = np.maximum.accumulate(np.concatenate(([mx], flat)))[:-1]
    raises = flat > run_max
    mx = max(mx, flat.max()) if flat.size else mx
    cand = flat[~raises]
    if cand.size:
        mn = min(mn, cand.min())
    scaled = (flat - mn) * np.float32(255.0) / np.float32(mx - mn)
    return scaled.astype(np.int32).astype(np.uint8).reshape(mat.shape)


def run(left_rgb, right_rgb, dmin=-15, dmax=0, fused=False, device=0, subpixel=False):
    """returns {file name: uint8 image} (and, with subpixel, one float32 array under a .npy name)"""
    fused = fused or subpixel
    p = default_params(dmin=dmin, dmax=dmax, box_mode=BOX_SAT)
    init = np.frombuffer(np.array([0x7F7F7F7F], np.uint32).tobytes(), np.float32)[0]  # main.cu:112
    with Context(device) as ctx:
        gl, gr = ctx.rgb_to_grayscale(left_rgb, p), ctx.rgb_to_grayscale(right_rgb, p)
        costl = ctx.compute_cost(gl, gr, dmin, p)          # main.cu:80
        costr = ctx.compute_cost(gr, gl, -dmax, p)         # main.cu:82
        if fused:
            want = ctx._F32 + ctx._U8 + (("subpixel_left",) if subpixel else ())
            out = ctx.pipeline(gl, gr, p, want=want)
            bl, br, dl, dr = out["best_left"], out["best_right"], out["disp_left"], out["disp_right"]
            ml, mr, occ, filled = out["mean_left"], out["mean_right"], out["occlusion"], out["filled"]
        else:
            bl, br = np.full(gl.shape, init, np.float32), np.full(gl.shape, init, np.float32)
            dl, dr = np.zeros(gl.shape, np.float32), np.zeros(gl.shape, np.float32)
            ml = ctx.compute_guided_filter(gl, costl, bl, dl, dmin, p)     # main.cu:133
            mr = ctx.compute_guided_filter(gr, costr, br, dr, -dmax, p)    # main.cu:134
            occ = dl.copy()
            ctx.detect_occlusion(occ, dr, dmin - 100, p)                   # main.cu:149-150
            filled = occ.copy()
            ctx.fill_occlusion(filled, dmin)                               # main.cu:153-155
        # main.cu:162-181: the float maps are normalised to 8 bits on the device (sb200_write_mat)
        wm = ctx.write_mat
        tag = f"minus{-dmin}" if dmin < 0 else str(dmin)
        extra = {}
        if subpixel:
            extra = {"disparity_mapl_subpixel.png": wm(out["subpixel_left"]), "disparity_mapl_subpixel.npy": out["subpixel_left"]}
        return {
            **extra,
            "image_left.png": gl, "image_right.png": gr, "image_mean_left.png": ml, "image_mean_right.png": mr,
            "best_costl.png": wm(bl), "best_costr.png": wm(br),
            f"cost_l{tag}.png": wm(costl[0]), f"cost_r{tag}.png": wm(costr[0]),
            "occlu_mapl.png": wm(occ), "disparity_mapl.png": wm(dl), "disparity_mapr.png": wm(dr),
            "occlu_mapl_filled.png": wm(filled),
        }


def main():
    from PIL import Image

    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("left")
    ap.add_argument("right")
    ap.add_argument("outdir")
    ap.add_argument("--dmin", type=int, default=-15)
    ap.add_argument("--dmax", type=int, default=0)
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--subpixel", action="store_true", help="also write the sub-pixel refined left disparity (beyond the reference)")
    a = ap.parse_args()
    L = np.array(Image.open(a.left).convert("RGB"))
    R = np.array(Image.open(a.right).convert("RGB"))
    os.makedirs(a.outdir, exist_ok=True)
    for name, img in run(L, R, a.dmin, a.dmax, a.fused, subpixel=a.subpixel).items():
        if name.endswith(".npy"):
            np.save(os.path.join(a.outdir, name), img)
        else:
            Image.fromarray(img).save(os.path.join(a.outdir, name))
        print("wrote", os.path.join(a.outdir, name))


if __name__ == "__main__":
    main()
