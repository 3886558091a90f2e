#!/usr/bin/env python
"""Does a sustained loop of the fused pipeline run at the clock a short burst runs at?  Loops the RGB-guide (or gray) pipeline
on device-resident 1080p D=256 pairs for a few seconds while nvidia-smi samples SM clock, power draw and throttle reasons,
and prints ms/pair per 50-pair window:  python tools/power_probe.py [gray|rgb] [seconds]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
import stereo_matching_cuda_b200 as S
from stereo_matching_cuda_b200 import api

guide = sys.argv[1] if len(sys.argv) > 1 else "rgb"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
w, h, D = 1920, 1080, 256
ch = 3 if guide == "rgb" else 1
L, R = synth.make_pair(w, h, D, channels=ch, seed=3) if ch == 3 else synth.make_pair(w, h, D, seed=3)
dev = torch.device("cuda:0")
dl, dr = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
outs = {k: torch.empty((h, w), dtype=torch.float32, device=dev) for k in ("disp_left", "disp_right", "occlusion", "filled")}
p = api.default_params(dmin=-(D - 1), dmax=0, guide_mode=S.GUIDE_RGB if ch == 3 else S.GUIDE_GRAY)
with S.Context(0) as ctx:
    for _ in range(3):
        ctx.pipeline_dev(dl, dr, ch, w, h, outs, p)
    torch.cuda.synchronize()
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader", "-lms", "100", "-i", "0"],
                           stdout=subprocess.PIPE, text=True)
    t_end = time.time() + secs
    win = []
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ctx.pipeline_dev(dl, dr, ch, w, h, outs, p)
        e1.record(); torch.cuda.synchronize()
        win.append(e0.elapsed_time(e1) / 50)
    smi.terminate()
    print(guide, "ms/pair per 50-pair window:", " ".join(f"{x:.3f}" for x in win))
    print("nvidia-smi samples (MHz, W, reasons):")
    for line in smi.stdout.read().strip().splitlines()[:60]:
        print("  ", line)
