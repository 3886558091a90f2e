#!/usr/bin/env python
"""bench.py -- pixel.disparities/sec and fps of the local stereo pipeline on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c5|c1]

A "step" is one stereo pair through the whole hot path: rgb->gray (when the input has
colour), guide statistics, fused cost + guided filter + WTA for BOTH views, L/R check and
fill.  Default workload = BASELINE.json configs[2] shape: synthetic 1920x1080, D=256.
`value` counts 2*W*H*D cells per pair (both views) with inputs resident in HBM; `e2e` is
the same metric through the host-pointer C-ABI call (sb200_pipeline) with pinned host
buffers, H2D and D2H inside the timed region.

N > 1 (launched by torchrun): one process per GPU, pairs are data-parallel with no
communication, weak scaling: every rank runs the same per-rank work.

--impl reference times the reference's own CPU implementation (oracle/_ref: its unmodified
functions, chained in main.cu order) on the host cores, on a bounded sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

# workload table: name -> (w, h, size_d, channels, description)
WORKLOADS = {
    "c1": (384, 288, 16, 3, "Tsukuba-shaped synthetic pair 384x288 RGB, D=16"),
    "c2": (3052, 1968, 16, 3, "bike-shaped synthetic pair 3052x1968 RGB, D=16"),
    "c3": (1920, 1080, 256, 1, "synthetic 1920x1080 rectified pair (hash texture + 16-band disparity staircase), D=256"),
    "c4": (1920, 1080, 128, 1, "synthetic 1920x1080 pairs, D=128 (batch data-parallel)"),
    "c5": (7680, 4320, 512, 1, "synthetic 7680x4320 pair, D=512"),
}
INSTR_PER_CELL = 35  # useful FP32 instructions per (pixel, disparity) cell, gray guide (RGB: 71)
FLOP_PER_CELL = 39


def peaks():
    p = {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(hbm_gbs=float(m["hbm_gbs"]), sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)), src="measured")
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 100 ms from before the warm-up until
    after the timed region; samples are stamped on arrival and only those inside the timed
    region (or, if it was shorter than the sampling period, inside the loaded window) count."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1, load0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def parse(lo, hi):
            sm, smax, reasons = [], [], set()
            for ts, ln in self.lines:
                if not (lo <= ts <= hi):
                    continue
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, smax, reasons

        sm, smax, reasons = parse(t0, t1 + 0.1)
        window = "timed region"
        if len(sm) < 3:  # timed region shorter than a few sampling periods: use the whole loaded window
            sm, smax, reasons = parse(load0, t1 + 0.1)
            window = "warm-up + timed region"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


class quiet_stdout:
    """the reference's detect_occlusionOnCPU prints to stdout (occlusion.cu:106); keep ours to one JSON line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def make_inputs(w, h, size_d, channels, n_sets, y0=0, rows=None, seed0=0):
    import synth

    return [synth.make_pair(w, h, size_d, channels=channels, seed=seed0 + s, y0=y0, rows=rows) for s in range(n_sets)]


def ncu_traffic(guide="gray"):
    """DRAM bytes per launch (and pipe utilisations) of the dominant kernel, from the committed ncu --set full capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "fused_ncu.json" if guide == "gray" else "fused_rgb3_ncu.json")) as f:
            return json.load(f)
    except Exception:
        return None


def run_reference(args, w, h, size_d, desc):
    """CPU arm: the reference's own functions (oracle/_ref) or the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import _oracle as O

    threads = os.cpu_count() or 1
    if O.ref_available():
        lib, kind = O.load_ref(), "reference"
        threads = min(threads, lib.max_threads())
    else:
        lib, kind = O.load_oracle(), "port"
        threads = min(threads, lib.max_threads())
    d_sample = int(min(size_d, max(8, 2 * threads)))
    dmin_full = -(size_d - 1)
    L, R = make_inputs(w, h, size_d, 1, 1)[0]
    init = O.load_oracle().best_init()

    def step():
        # a contiguous block of the disparity range, both views, then L/R check + fill
        if kind == "reference":
            bl, dl, _ = lib.view_disparity_cpu(L, R, d_sample, dmin_full, init, nthreads=threads)
            br, dr, _ = lib.view_disparity_cpu(R, L, d_sample, 0, init, nthreads=threads)
            with quiet_stdout():
                occ = lib.detect_occlusion_cpu(dl, dr, dmin_full - 100)
            lib.fill_occlusion_cpu(occ, dmin_full)
        else:
            p = lib.params(box_mode=O.BOX_FAITHFUL, nthreads=threads)
            bl, dl, _, _ = lib.view_disparity(L, R, d_sample, dmin_full, p)
            br, dr, _, _ = lib.view_disparity(R, L, d_sample, 0, p)
            lib.fill_occlusion(lib.detect_occlusion(dl, dr, dmin_full - 100), dmin_full)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    cells = 2.0 * w * h * d_sample * args.steps
    value = cells / dt
    sample = f"{w}x{h}, {d_sample} of {size_d} disparities (d={dmin_full}..{dmin_full + d_sample - 1}), both views + L/R check + fill"
    line = {
        "impl": "reference", "metric": "pixel-disparities/sec", "value": value, "unit": "px*d/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc + ", gray guide, r=9, eps=6.5025; CPU arm runs a bounded sample: " + sample},
        "cpu_baseline": {"value": value, "unit": "px*d/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "px*d/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps_equiv_full_workload": value / (2.0 * w * h * size_d),
    }
    print(json.dumps(line))


def cpu_baseline(w, h, size_d):
    """single-threaded reference CPU chain on a bounded sample (rank 0, N=1 only)"""
    import _oracle as O

    L, R = make_inputs(w, h, size_d, 1, 1)[0]
    dmin_full = -(size_d - 1)
    # ~10-20 s of single-thread work at ~14 M cells/s
    d_sample = max(1, min(size_d, int(180e6 / (2.0 * w * h))))
    init = O.load_oracle().best_init()
    if O.ref_available():
        lib, kind = O.load_ref(), "reference"
        t0 = time.perf_counter()
        _, dl, _ = lib.view_disparity_cpu(L, R, d_sample, dmin_full, init, nthreads=1)
        _, dr, _ = lib.view_disparity_cpu(R, L, d_sample, 0, init, nthreads=1)
        with quiet_stdout():
            occ = lib.detect_occlusion_cpu(dl, dr, dmin_full - 100)
        lib.fill_occlusion_cpu(occ, dmin_full)
        dt = time.perf_counter() - t0
    else:
        lib, kind = O.load_oracle(), "port"
        p = lib.params(box_mode=O.BOX_FAITHFUL, nthreads=1)
        t0 = time.perf_counter()
        _, dl, _, _ = lib.view_disparity(L, R, d_sample, dmin_full, p)
        _, dr, _, _ = lib.view_disparity(R, L, d_sample, 0, p)
        lib.fill_occlusion(lib.detect_occlusion(dl, dr, dmin_full - 100), dmin_full)
        dt = time.perf_counter() - t0
    return {"value": 2.0 * w * h * d_sample / dt, "unit": "px*d/s", "cores": 1, "kind": kind,
            "sample": f"{w}x{h}, {d_sample} of {size_d} disparities, both views + L/R check + fill, {dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SB200_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default=None, choices=["dp", "batch", "strips"],
                    help="dp: one pair per step per GPU, weak scaling (default); batch: a fixed batch of 64 pairs "
                         "split over the GPUs (c4); strips: one frame split into row strips with halo exchange (c5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--guide", default="gray", choices=["gray", "rgb"],
                    help="gray: the reference's only guide mode (fused kernel, parity pinned to its goldens); rgb: colour "
                         "guided filter of SURVEY A.8, which the reference does not have (staged path, parity unpinned)")
    args = ap.parse_args()
    w, h, size_d, channels, desc = WORKLOADS[args.workload]
    if args.guide == "rgb":
        channels = 3

    if args.impl == "reference":
        run_reference(args, w, h, size_d, desc)
        return

    import ctypes as C

    import torch
    import torch.distributed as dist

    import stereo_matching_cuda_b200 as S
    from stereo_matching_cuda_b200 import api, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL_DEBUG=VERSION/INFO makes NCCL print to STDOUT, which would break the one-JSON-line contract
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = os.path.join("/tmp", "nccl_bench_%h_%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = args.mode or {"c4": "batch", "c5": "strips" if world > 1 else "dp"}.get(args.workload, "dp")

    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB if args.guide == "rgb" else S.GUIDE_GRAY)
    n = w * h
    dev = torch.device("cuda", local)
    ctx = S.Context(local, stream=torch.cuda.current_stream())
    names = ("disp_left", "disp_right", "occlusion", "filled", "best_left", "best_right")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    halo = ctx.strip_halo_rows(p)
    if mode == "strips":
        # one frame, rows split over the ranks; every step exchanges the 2*radius input halo rows
        geom = sharding.strip_geometry(h, rank, world, halo)
        n_sets = 2
        own = make_inputs(w, h, size_d, channels, n_sets, y0=geom["y0"], rows=geom["rows"])
        d_own = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in own]
        outs = {k: torch.empty((geom["rows"], w), dtype=torch.float32, device=dev) for k in names}
        pairs_per_step, cells_per_step, scaling = 1, 2.0 * w * h * size_d, "strong"

        def step(i):
            a, b = d_own[i % n_sets]
            if world > 1:
                a = sharding.exchange_halo_rows(a, geom, rank, world, halo)
                b = sharding.exchange_halo_rows(b, geom, rank, world, halo)
            ctx.pipeline_strip_dev(a, b, channels, w, geom, outs, p)
    else:
        n_sets = 4 if n * size_d < 3e9 else 2
        if mode == "batch":
            total_pairs = 64
            mine = sharding.batch_shard(total_pairs, rank, world)
            n_sets = min(4, len(mine))
            pairs = [make_inputs(w, h, size_d, channels, 1, seed0=mine[k])[0] for k in range(n_sets)]
            pairs_per_step, cells_per_step, scaling = len(mine), 2.0 * w * h * size_d * total_pairs, "strong"
        else:
            pairs = make_inputs(w, h, size_d, channels, n_sets)
            pairs_per_step, cells_per_step, scaling = 1, 2.0 * w * h * size_d * world, "weak"
        d_in = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in pairs]
        outs = {k: torch.empty((h, w), dtype=torch.float32, device=dev) for k in names}

        def step(i):
            for k in range(pairs_per_step):
                a, b = d_in[(i * pairs_per_step + k) % n_sets]
                ctx.pipeline_dev(a, b, channels, w, h, outs, p)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    load0 = time.time()
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - l0
    clocks = sampler.stop(t0, t1, load0) if rank == 0 else None
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
    tl = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    ms_max = float(t[0].item())
    launches_all = int(tl[1].item())
    value = cells_per_step * args.steps / (ms_max * 1e-3)
    pairs_total_per_step = {"dp": world, "batch": 64, "strips": 1}[mode]

    # --- roofline pass: device time of the dominant kernel (k_fused_cvf) from CUDA events the
    # library records on the launching stream around that kernel, averaged over up to 10 steps
    ctx.enable_timing(True)
    fused_ms, occl_ms, prep_ms, merge_ms = [], [], [], []
    for i in range(min(args.steps, 10)):
        step(i)
        tm = ctx.last_timing()  # the step's last pair
        fused_ms.append(tm["fused_ms"])
        occl_ms.append(tm["occl_ms"])
        prep_ms.append(tm["prep_ms"])
        merge_ms.append(tm["merge_ms"])
    ctx.enable_timing(False)
    fk = statistics.mean(fused_ms)
    rows_local = h if mode != "strips" else sharding.strip_geometry(h, rank, world, halo)["rows"]
    cells_per_launch = 2.0 * w * rows_local * size_d

    # --- end to end through the host-pointer C-ABI call, pinned host buffers, copies timed
    e2e = None
    if mode == "dp":
        h_in = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in pairs]
        onames = ("disp_left", "disp_right", "occlusion", "filled")
        h_out = {k: torch.empty((h, w), dtype=torch.float32).pin_memory() for k in onames}
        o = api._Outputs()
        for k in onames:
            setattr(o, k, h_out[k].data_ptr())

        def e2e_step(i):
            a, b = h_in[i % n_sets]
            ctx._ck(ctx.lib.sb200_pipeline(ctx.h, C.byref(p), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), channels,
                                           w, h, C.byref(o)))

        for i in range(2):
            e2e_step(i)
        barrier()
        tt0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 10))
        for i in range(n_e2e):
            e2e_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - tt0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * 2.0 * w * h * size_d * n_e2e / float(te.item()), "unit": "px*d/s",
               "h2d_bytes_per_step": int(2 * n * channels), "d2h_bytes_per_step": int(4 * n * 4),
               "ms_per_step": 1e3 * float(te.item()) / n_e2e,
               "api": "sb200_pipeline (host pointers, pinned, blocking; H2D of the pair and D2H of 4 float maps inside)"}

    # --- the same shape with the RGB guide BASELINE configs[2] names (the reference has no such
    # mode, so it is reported beside the gray-guide headline rather than instead of it)
    rgb_extra = None
    if mode == "dp" and args.guide == "gray" and args.workload == "c3":
        p_rgb = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
        cpairs = make_inputs(w, h, size_d, 3, 2)
        c_in = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in cpairs]
        for i in range(3):
            ctx.pipeline_dev(c_in[i % 2][0], c_in[i % 2][1], 3, w, h, outs, p_rgb)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rgb = max(3, min(args.steps, 10))
        r0.record()
        for i in range(n_rgb):
            ctx.pipeline_dev(c_in[i % 2][0], c_in[i % 2][1], 3, w, h, outs, p_rgb)
        r1.record()
        barrier()
        tr_ms = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr_ms, op=dist.ReduceOp.MAX)
        ctx.enable_timing(True)
        ctx.pipeline_dev(c_in[0][0], c_in[0][1], 3, w, h, outs, p_rgb)
        k_rgb = ctx.last_timing()["fused_ms"]
        ctx.enable_timing(False)
        ms_rgb = float(tr_ms.item()) / n_rgb
        rgb_extra = {"value": world * 2.0 * w * h * size_d / (ms_rgb * 1e-3), "unit": "px*d/s", "ms_per_step": ms_rgb,
                     "fps": world / (ms_rgb * 1e-3), "kernel": "k_fused_cvf_rgb3" if ctx.rgb_kernel == 3 else "k_fused_cvf_rgb", "kernel_ms": k_rgb, "instr_per_cell": 71,
                     "roofline_frac": 71 * 2.0 * w * h * size_d / (k_rgb * 1e-3) / (148 * 128 * peaks()["sm_max_mhz"] * 1e6),
                     "note": "colour guided filter of SURVEY A.8 on 3-channel synthetic pairs; not in the reference "
                             "(parity unpinned; checked against the oracle's RGB port and the eps/3 identity)"}

    if rank == 0:
        pk = peaks()
        peak_instr = 148 * 128 * pk["sm_max_mhz"] * 1e6  # FP32 lane-instructions/s at the max SM clock
        ipc = INSTR_PER_CELL if args.guide == "gray" else 71
        achieved = ipc * cells_per_launch / (fk * 1e-3)
        tr = ncu_traffic(args.guide)
        line = {
            "metric": "pixel-disparities/sec", "value": value, "unit": "px*d/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "fps": pairs_total_per_step * args.steps / (ms_max * 1e-3),
            "config": {
                "workload": desc + f" (dmin={-(size_d - 1)}), " + ("gray guide (the reference's only guide mode; BASELINE configs[2] names an RGB guide, which the reference does not implement: see --guide rgb)" if args.guide == "gray" else "RGB guide (SURVEY A.8; not in the reference)") + ", r=9, eps=6.5025, both views + L/R check + fill; "
                            + {"dp": "1 pair per step per GPU" + ("" if world == 1 else f", {world} GPUs data-parallel over pairs, no communication"),
                               "batch": f"a batch of 64 pairs per step split over {world} GPU(s), no communication",
                               "strips": f"one frame per step split into {world} row strips, {halo}-row input halos exchanged over NCCL send/recv"}[mode],
                "mode": mode, "width": w, "height": h, "size_d": size_d, "channels": channels,
                "l2": f"inputs rotate over {n_sets} distinct pairs; the per-pair working set (prepared planes + per-chunk "
                      "WTA planes, several hundred MB at 1080p D=256) exceeds the 126 MB L2",
            },
            "roofline": {
                "bound": "fp32_pipe", "kernel": "k_fused_cvf" if args.guide == "gray" else ("k_fused_cvf_rgb3" if ctx.rgb_kernel == 3 else "k_fused_cvf_rgb"), "achieved": achieved / 1e12, "peak": peak_instr / 1e12,
                "unit": "T lane-instr/s", "frac": achieved / peak_instr,
                "traffic": tr.get("dram_bytes_per_launch") if tr and mode == "dp" and args.workload == "c3" else None,
                "traffic_src": tr.get("src") if tr and mode == "dp" and args.workload == "c3" else None,
                "kernel_ms": fk, "instr_per_cell": ipc,
                "flop_frac": (FLOP_PER_CELL if args.guide == "gray" else 87) * cells_per_launch / (fk * 1e-3) / (2 * peak_instr),
                "peak_src": f"148 SMs x 128 FP32 lanes x {pk['sm_max_mhz']:.0f} MHz (sm_max_mhz, {pk['src']})",
                "note": "north_star names the FP32 CUDA-core pipe as this kernel's roofline (no dense contraction; see "
                        "hbm_frac_of_kernel_time for how little of the kernel's time its DRAM traffic explains)",
                "hbm_frac_of_kernel_time": (tr["dram_bytes_per_launch"] / (pk["hbm_gbs"] * 1e9)) / (fk * 1e-3) if tr and mode == "dp" and args.workload == "c3" else None,
                "ncu": {k: tr.get(k) for k in ("l1tex_data_pipe_pct", "issue_active_pct", "fma_pipe_pct", "alu_pipe_pct",
                                               "kernel_ms_under_ncu")} if tr and mode == "dp" and args.workload == "c3" else None,
                "ncu_note": "the kernel's busiest unit is the L1TEX/shared-memory data pipe (warp shuffles of the horizontal window "
                            "sums + 128-bit shared loads of the operand ring), not the FP32 pipe: l1tex_data_pipe_pct is its "
                            "utilisation in the committed ncu capture",
                "other_kernels_ms": {"k_prep_x2": statistics.mean(prep_ms), "k_merge_chunks_x2": statistics.mean(merge_ms),
                                     "k_lr_check_fill": statistics.mean(occl_ms)},
                "lr_check_fill_hbm": {"bound": "hbm", "achieved": 16.0 * w * rows_local / (statistics.mean(occl_ms) * 1e-3) / 1e9,
                                      "peak": pk["hbm_gbs"], "unit": "GB/s", "bytes_per_pixel": 16},
            },
            "gpu_launches": launches_all,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if rgb_extra:
            line["rgb_guide"] = rgb_extra
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(w, h, size_d)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
