"""Times the reference's OWN GPU path (oracle/_ref/libref.so: its unmodified .cu files recompiled for sm_100a) on a
synthetic RGB pair of the given size, host buffers in and out, in main.cu's call order:
    python tools/time_ref_gpu.py W H
Prints one JSON line {"ms_per_pair": ...}.  Child process of bench.py's `reference_gpu` leg (a fault or hang in the
reference's kernels must not take the bench down).  D and dmin are the reference's compile-time macros (16, -15)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as O  # noqa: E402
import synth  # noqa: E402


def pair(ref, L, R):
    h, w, _ = L.shape
    n = h * w
    gl, gr = np.empty((h, w), np.uint8), np.empty((h, w), np.uint8)
    ref.lib.ref_rgb_to_gray_gpu(L, gl, n, 3)
    ref.lib.ref_rgb_to_gray_gpu(R, gr, n, 3)
    size_d, dmin = ref.size_d_macro, ref.dmin_macro
    cost_l = np.zeros((size_d, h, w), np.float32)
    cost_r = np.zeros((size_d, h, w), np.float32)
    ref.lib.ref_compute_cost_gpu(gl, gr, cost_l, w, h, dmin)
    ref.lib.ref_compute_cost_gpu(gr, gl, cost_r, w, h, 0)
    init = np.frombuffer(np.array([0x7F7F7F7F], np.uint32).tobytes(), np.float32)[0]
    maps = []
    for img, cost, dm in ((gl, cost_l, dmin), (gr, cost_r, 0)):
        best = np.full((h, w), init, np.float32)
        dmap = np.zeros((h, w), np.float32)
        mean = np.zeros((h, w), np.uint8)
        ref.lib.ref_compute_guided_filter_gpu(img, cost, best, dmap, mean, w, h, size_d, dm)
        maps.append(dmap)
    occ = maps[0].copy()
    ref.lib.ref_detect_occlusion_gpu(occ, maps[1], dmin - 100, w, h)
    ref.lib.ref_fill_occlusion_gpu(occ, w, h, float(dmin))
    return occ


def main():
    w, h = int(sys.argv[1]), int(sys.argv[2])
    ref = O.load_ref()
    L, R = synth.make_pair(w, h, 16, channels=3, seed=0)
    pair(ref, L, R)  # warm-up (context, module load)
    reps = 2 if w * h > 10 ** 6 else 5
    t0 = time.perf_counter()
    for _ in range(reps):
        pair(ref, L, R)
    ms = 1e3 * (time.perf_counter() - t0) / reps
    print(json.dumps({"ms_per_pair": ms, "reps": reps, "shape": [w, h], "size_d": int(ref.size_d_macro)}))


if __name__ == "__main__":
    main()
