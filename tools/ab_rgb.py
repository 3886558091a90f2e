#!/usr/bin/env python
"""A/B the fused RGB-guide kernels on one box, no torch needed (host-pointer pipeline + the library's own CUDA-event
timing of the fused kernel):
    python tools/ab_rgb.py [name=path/to/lib.so[:kernel] ...]
AB_GUIDE=gray times the gray-guide kernel k_fused_cvf instead.  Each variant runs in a child process (SB200_LIB / SB200_RGB_KERNEL are read at load / context creation).  Prints the
fused-kernel ms (min and median of 5 runs at 1920x1080 D=256) and a CRC of the outputs (equal CRCs = bit-identical)."""
import json
import os
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(npz):
    import stereo_matching_cuda_b200 as S
    from stereo_matching_cuda_b200 import api

    d = np.load(npz)
    L, R, size_d = d["L"], d["R"], int(d["size_d"])
    gray = os.environ.get("AB_GUIDE", "rgb") == "gray"
    if gray:  # the gray-guide kernel on the gray versions of the same pair
        with S.Context(0) as c0:
            L, R = c0.rgb_to_grayscale(L), c0.rgb_to_grayscale(R)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_GRAY if gray else S.GUIDE_RGB)
    with S.Context(0) as ctx:
        ctx.enable_timing()
        ms = []
        for _ in range(6):
            out = ctx.pipeline(L, R, p, want=("disp_left", "disp_right", "best_left", "best_right"))
            ms.append(ctx.last_timing()["fused_ms"])
        crc = 0
        for k in sorted(out):
            crc = zlib.crc32(out[k].tobytes(), crc)
        print(json.dumps({"kernel": ctx.rgb_kernel, "fused_ms_min": min(ms[1:]), "fused_ms_med": float(np.median(ms[1:])),
                          "crc": crc}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    import synth

    w, h, size_d = 1920, 1080, 256
    L, R = synth.make_pair(w, h, size_d, channels=3, seed=3)
    npz = "/tmp/ab_rgb_pair.npz"
    np.savez(npz, L=L, R=R, size_d=size_d)
    for spec in sys.argv[1:] or ["head=" + os.path.join(ROOT, "stereo_matching_cuda_b200", "libstereo_b200.so") + ":2"]:
        name, rest = spec.split("=", 1)
        path, _, kern = rest.partition(":")
        env = dict(os.environ, SB200_LIB=os.path.abspath(path))
        if kern:
            env["SB200_RGB_KERNEL"] = kern
        try:
            r = subprocess.run([sys.executable, __file__, "--child", npz], env=env, capture_output=True, text=True, timeout=45)
            print(name, r.stdout.strip() or ("FAILED: " + r.stderr.strip()[-400:]), flush=True)
        except subprocess.TimeoutExpired:
            print(name, "TIMEOUT (45 s)", flush=True)
