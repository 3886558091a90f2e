#!/usr/bin/env python
"""k_lr_check_fill alone (L/R check + scan-line fill, occlusion.cu:3-15 and :111-176): time and HBM rate at 8K and 1080p on
random label maps (16 B per pixel: two maps read, two written).  python tools/lr_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stereo_matching_cuda_b200 as S  # noqa: E402
from stereo_matching_cuda_b200 import api  # noqa: E402

dev = torch.device("cuda:0")
for (w, h, D) in ((7680, 4320, 512), (1920, 1080, 256)):
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    dL = -torch.randint(0, D, (h, w), device=dev, generator=g).float()
    dR = torch.randint(0, D, (h, w), device=dev, generator=g).float()
    occ, fil = torch.empty_like(dL), torch.empty_like(dL)
    p = api.default_params(dmin=-(D - 1), dmax=0)
    with S.Context(0) as ctx:
        for _ in range(3):
            ctx.lr_check_fill_dev(dL, dR, w, h, -(D - 1) - 100, float(-(D - 1)), occ, fil, params=p)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            ctx.lr_check_fill_dev(dL, dR, w, h, -(D - 1) - 100, float(-(D - 1)), occ, fil, params=p)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{w}x{h}: {ms:.4f} ms  {16.0 * w * h / ms / 1e6:.0f} GB/s of 6468 measured peak = {16.0 * w * h / ms / 1e6 / 6468:.2f}")
