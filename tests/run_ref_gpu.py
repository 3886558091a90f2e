"""Runs the reference's OWN GPU host functions (oracle/_ref/libref.so: its unmodified .cu files
recompiled for sm_100a) on the Tsukuba pair and stores the results in gpurun_out/.  Executed as a
child process by tests/test_gpu_parity.py so that a fault in the reference's kernels cannot take
the test session down.  GPU box only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _oracle as O  # noqa: E402


def main():
    ref = O.load_ref()
    L, R = O.tsukuba_rgb()
    h, w, _ = L.shape
    n = h * w
    gl, gr = np.empty((h, w), np.uint8), np.empty((h, w), np.uint8)
    ref.lib.ref_rgb_to_gray_gpu(np.ascontiguousarray(L), gl, n, 3)
    ref.lib.ref_rgb_to_gray_gpu(np.ascontiguousarray(R), gr, n, 3)
    size_d, dmin = ref.size_d_macro, ref.dmin_macro
    cost_l = np.zeros((size_d, h, w), np.float32)
    cost_r = np.zeros((size_d, h, w), np.float32)
    ref.lib.ref_compute_cost_gpu(gl, gr, cost_l, w, h, dmin)
    ref.lib.ref_compute_cost_gpu(gr, gl, cost_r, w, h, 0)
    init = np.frombuffer(np.array([0x7F7F7F7F], np.uint32).tobytes(), np.float32)[0]
    res = {"gray_l": gl, "gray_r": gr, "cost_l": cost_l}
    for name, img, cost, dm in (("l", gl, cost_l, dmin), ("r", gr, cost_r, 0)):
        best = np.full((h, w), init, np.float32)
        dmap = np.zeros((h, w), np.float32)
        mean = np.zeros((h, w), np.uint8)
        ref.lib.ref_compute_guided_filter_gpu(img, cost, best, dmap, mean, w, h, size_d, dm)
        res["best_" + name], res["dmap_" + name], res["mean_" + name] = best, dmap, mean
    occ = res["dmap_l"].copy()
    ref.lib.ref_detect_occlusion_gpu(occ, res["dmap_r"], dmin - 100, w, h)
    filled = occ.copy()
    ref.lib.ref_fill_occlusion_gpu(filled, w, h, float(dmin))
    res["occ"], res["filled"] = occ, filled
    out = os.path.join(HERE, "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "ref_gpu_tsukuba.npz"), **res)
    print("reference GPU path ok")


if __name__ == "__main__":
    main()
