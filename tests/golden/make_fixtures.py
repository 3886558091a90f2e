#!/usr/bin/env python
"""Regenerates tests/golden/ from the read-only reference checkout.

Run in the authoring container (needs /root/reference and a built oracle/_ref/libref.so):

    make -C oracle && python tests/golden/make_fixtures.py

1. tsukuba/: the reference's own input pair and the 12 PNGs its main() wrote
   (stereo_matching_cuda/data/, main.cu:162-181).  The two stale "... - Copie.png"
   files are skipped.  These are data fixtures, not sources.
2. ref_twins_small.npz: outputs of the reference's CPU twins (oracle/_ref/libref.so,
   compiled from the unmodified sources) on small seeded inputs, so the oracle can be
   pinned against the reference's code on machines where /root/reference is absent.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF_DATA = "/root/reference/stereo_matching_cuda/data"

NAMES = [
    "tsukuba0.png", "tsukuba1.png", "image_left.png", "image_right.png", "image_mean_left.png",
    "image_mean_right.png", "cost_lminus15.png", "cost_rminus15.png", "best_costl.png", "best_costr.png",
    "disparity_mapl.png", "disparity_mapr.png", "occlu_mapl.png", "occlu_mapl_filled.png",
]


def main():
    out = os.path.join(HERE, "tsukuba")
    os.makedirs(out, exist_ok=True)
    for n in NAMES:
        shutil.copyfile(os.path.join(REF_DATA, n), os.path.join(out, n))
    import _oracle as O

    ref = O.load_ref()
    rng = np.random.default_rng(1234)
    w, h, size_d, dmin = 61, 47, 5, -3
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    base = rng.integers(0, 256, size=(h, w + 8), dtype=np.uint8)
    # smooth a little so costs are not all saturated
    base = ((base.astype(np.int32) + np.roll(base, 1, 1) + np.roll(base, -1, 1) + np.roll(base, 1, 0)) // 4).astype(np.uint8)
    right = np.ascontiguousarray(base[:, 4:4 + w])
    left = np.ascontiguousarray(base[:, 2:2 + w])
    res = {"w": w, "h": h, "size_d": size_d, "dmin": dmin, "rgb": rgb, "left": left, "right": right}
    res["gray"] = ref.rgb_to_gray_cpu(rgb)
    res["grad_left"] = ref.x_derivative_cpu(left)
    res["cost"] = ref.cost_volume_cpu(left, right, size_d, dmin)
    fimg = left.astype(np.float32) * 1.5
    res["sat"] = ref.integral_cpu(fimg)
    res["box"] = ref.box_filter_cpu(fimg, res["sat"])
    dL, dR, occ, filled, bL, bR = ref.pipeline_gray_cpu(left, right, dmin, size_d, nthreads=1)
    res.update(dL=dL, dR=dR, occ=occ, filled=filled, bestL=bL, bestR=bR)
    np.savez_compressed(os.path.join(HERE, "ref_twins_small.npz"), **res)
    print("wrote", out, "and ref_twins_small.npz")


if __name__ == "__main__":
    main()
