"""GPU parity tests: every entry point of the C ABI against the CPU oracle on the same inputs.

Bars (north_star): bit-exact for integer/byte/index work and for every stage whose float32
arithmetic is specified operation by operation (gray, gradient, cost, float32 SAT, 4-tap box,
SAT-mode guided filter, WTA, occlusion check, fill); the fused sliding-window path is compared
with the oracle's EXACT box mode: best cost within 1e-4 relative, labels identical wherever the
oracle's WTA margin exceeds the float tolerance and >= 99.9 % overall.
"""
import os

import numpy as np
import pytest

import _oracle as O
import synth

pytestmark = pytest.mark.gpu

S = pytest.importorskip("stereo_matching_cuda_b200")
from stereo_matching_cuda_b200 import api  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    c = S.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def tsukuba(oracle):
    L, R = O.tsukuba_rgb()
    return L, R, oracle.rgb_to_gray(L), oracle.rgb_to_gray(R)


RTOL_BEST = 1e-4  # north_star: intermediate cost and filter outputs within 1e-4 relative


def check_fused_vs_oracle(out, ref, margin_tau, min_agree=0.999):
    for lab, best, second, o_lab, o_best in (("dL", "bestL", "secondL", "disp_left", "best_left"),
                                             ("dR", "bestR", "secondR", "disp_right", "best_right")):
        same = out[o_lab] == ref[lab]
        assert same.mean() >= min_agree, f"{o_lab}: raw agreement {same.mean()}"
        margin = ref[second] - ref[best]
        decisive = margin > margin_tau
        assert np.all(same[decisive]), f"{o_lab}: {np.sum(~same & decisive)} decisive pixels differ"
        rel = np.abs(out[o_best] - ref[best]) / np.maximum(np.abs(ref[best]), 1e-2)
        assert rel.max() < RTOL_BEST, f"{o_best}: max rel err {rel.max()}"


# ---- stage drop-ins: bit-exact -------------------------------------------------------------
def test_rgb_to_grayscale_exact(ctx, oracle, tsukuba):
    L, R, gl, gr = tsukuba
    assert np.array_equal(ctx.rgb_to_grayscale(L), gl)
    rng = np.random.default_rng(0)
    rgba = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    assert np.array_equal(ctx.rgb_to_grayscale(rgba), oracle.rgb_to_gray(rgba))
    # every (r,g,b) whose exact luma is an integer is where double rounding could bite
    vals = np.array([(r, g, b) for r in range(0, 256, 5) for g in range(0, 256, 5) for b in range(0, 256, 5)], np.uint8)
    img = vals.reshape(1, -1, 3)
    assert np.array_equal(ctx.rgb_to_grayscale(img), oracle.rgb_to_gray(img))
    with pytest.raises(S.StereoB200Error):
        ctx.rgb_to_grayscale(np.zeros((4, 4, 1), np.uint8))


def test_x_derivative_and_cost_exact(ctx, oracle, tsukuba):
    _, _, gl, gr = tsukuba
    assert np.array_equal(ctx.x_derivative(gl), oracle.x_derivative(gl))
    assert np.array_equal(ctx.compute_cost(gl, gr, -15), oracle.cost_volume(gl, gr, 16, -15))
    assert np.array_equal(ctx.compute_cost(gr, gl, 0), oracle.cost_volume(gr, gl, 16, 0))
    # size_d > 64 (the reference's block shape cannot run it, costVolume.cu:39) and dmin > 0
    p = api.default_params(dmin=3, dmax=102)
    a, b = synth.make_pair(131, 17, 100)
    assert np.array_equal(ctx.compute_cost(a, b, 3, p), oracle.cost_volume(a, b, 100, 3))


def test_integral_and_sat_box_exact(ctx, oracle):
    rng = np.random.default_rng(1)
    for (w, h) in ((384, 288), (33, 70), (1, 5), (200, 1), (129, 129)):
        f = (rng.random((h, w), dtype=np.float32) * 255).astype(np.float32)
        sat = ctx.integral(f)
        assert np.array_equal(sat, oracle.integral(f)), (w, h)
        assert np.array_equal(ctx.box_filter_sat(sat), oracle.box_from_sat(sat)), (w, h)
    p = api.default_params(box_mode=S.BOX_SAT)
    f = (rng.random((97, 150), dtype=np.float32) * 1000).astype(np.float32)
    assert np.array_equal(ctx.box_filter(f, p), oracle.box_mean(f, oracle.params(box_mode=O.BOX_FAITHFUL)))


def test_sliding_box_matches_exact_oracle(ctx, oracle):
    rng = np.random.default_rng(2)
    for (w, h) in ((384, 288), (20, 15), (9, 40), (300, 7)):
        f = (rng.random((h, w), dtype=np.float32) * 637).astype(np.float32)
        got = ctx.box_filter(f, api.default_params(box_mode=S.BOX_SLIDING))
        want = oracle.box_mean(f, oracle.params(box_mode=O.BOX_EXACT))
        assert np.allclose(got, want, rtol=1e-6, atol=0), (w, h)


def test_filter_guide_statistics(ctx, oracle, tsukuba):
    _, _, gl, _ = tsukuba
    mean, var = ctx.filter(gl, api.default_params(box_mode=S.BOX_SAT))
    _, _, v, mu8 = oracle.guide_stats(gl, oracle.params(box_mode=O.BOX_FAITHFUL))
    assert np.array_equal(mean, mu8) and np.array_equal(var, v)
    assert np.array_equal(mean, O.load_png("image_mean_left.png"))


def test_guided_filter_sat_mode_bit_exact_and_goldens(ctx, oracle, tsukuba):
    """compute_guided_filter in SAT mode is the reference's arithmetic operation by operation:
    identical to the oracle's faithful mode, hence to the golden PNGs."""
    _, _, gl, gr = tsukuba
    p = api.default_params(box_mode=S.BOX_SAT)
    res = {}
    for name, (a, b, dmin) in {"l": (gl, gr, -15), "r": (gr, gl, 0)}.items():
        cost = ctx.compute_cost(a, b, dmin)
        best = np.full(a.shape, oracle.best_init(), np.float32)
        dmap = np.zeros(a.shape, np.float32)
        mean = ctx.compute_guided_filter(a, cost, best, dmap, dmin, p)
        ob, od, om, _ = oracle.guided_filter(a, cost, dmin, oracle.params(box_mode=O.BOX_FAITHFUL))
        assert np.array_equal(dmap, od) and np.array_equal(best, ob) and np.array_equal(mean, om)
        res[name] = (best, dmap)
    assert np.array_equal(res["l"][1].astype(int), O.load_png("disparity_mapl.png").astype(int) // 17 - 15)
    assert np.array_equal(res["r"][1].astype(int), O.load_png("disparity_mapr.png").astype(int) // 17)
    assert np.array_equal(oracle.write_mat(res["l"][0]), O.load_png("best_costl.png"))
    # occlusion + fill on those labels reproduce the remaining two goldens
    occ = res["l"][1].copy()
    ctx.detect_occlusion(occ, res["r"][1], -115)
    assert np.array_equal(occ == -115, O.load_png("occlu_mapl.png") == 0)
    ctx.fill_occlusion(occ, -15)
    assert np.array_equal(occ.astype(int), O.load_png("occlu_mapl_filled.png").astype(int) // 17 - 15)


def test_winner_take_all_tie_rule(ctx, oracle):
    rng = np.random.default_rng(3)
    n = 1000
    best = np.full(n, oracle.best_init(), np.float32)
    dmap = np.zeros(n, np.float32)
    q1 = rng.random(n, dtype=np.float32)
    ctx.winner_take_all(q1, best, dmap, -3)
    ctx.winner_take_all(q1, best, dmap, 4)  # identical slice: the LATER label must win (>=)
    assert np.all(dmap == 4) and np.array_equal(best, q1)
    q2 = q1 + np.where(rng.random(n) < 0.5, -0.1, 0.1).astype(np.float32)
    b2, d2 = best.copy(), dmap.copy()
    ctx.winner_take_all(q2, best, dmap, 9)
    oracle.disp_select(q2, b2, d2, 9)
    assert np.array_equal(best, b2) and np.array_equal(dmap, d2)


def test_occlusion_and_fill_exact(ctx, oracle):
    rng = np.random.default_rng(4)
    for (w, h, dmin) in ((384, 288, -15), (50, 3, -7), (1000, 2, -255), (17, 9, -3)):
        dL = rng.integers(dmin, 1, (h, w)).astype(np.float32)
        dR = rng.integers(0, -dmin + 1, (h, w)).astype(np.float32)
        occ = dL.copy()
        ctx.detect_occlusion(occ, dR, dmin - 100)
        want = oracle.detect_occlusion(dL, dR, dmin - 100)
        assert np.array_equal(occ, want)
        ctx.fill_occlusion(occ, dmin)
        assert np.array_equal(occ, oracle.fill_occlusion(want, dmin))
    # rows that are entirely occluded, or valid only at one end
    d = np.full((4, 40), -115, np.float32)
    d[1, 0] = -3
    d[2, -1] = -5
    d[3, 7] = -2
    d[3, 30] = -9
    got = d.copy()
    ctx.fill_occlusion(got, -15)
    assert np.array_equal(got, oracle.fill_occlusion(d, -15))
    assert np.all(got[0] == -15) and np.all(got[1] == -3) and np.all(got[3, 8:30] == -2)


def test_write_mat_and_fl_to_ch2_on_device(ctx, oracle):
    """main.cu:13-35 (write_mat) incl. its quirk -- a value that raises the running maximum is never a candidate for the
    minimum -- and occlusion.cu:230-237 (flToCh2OnGPU), bit-exact against the oracle / the formula."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    cases = [rng.random((37, 53), dtype=np.float32) * 100 - 30,
             np.arange(5000, dtype=np.float32).reshape(50, 100),           # every element raises the maximum: min stays at its init
             np.arange(5000, dtype=np.float32)[::-1].reshape(50, 100).copy(),
             rng.integers(-115, 1, (288, 384)).astype(np.float32),          # a disparity map with the occlusion sentinel
             rng.random((1, 1023), dtype=np.float32), rng.random((1, 1025), dtype=np.float32),
             (rng.random((1100, 1000), dtype=np.float32) - 0.5) * 7]        # more than 1024 blocks: the scan kernel loops
    ramp = np.sort(rng.random(3000).astype(np.float32))
    ramp[1500] = ramp[1499]                                                # a tie with the running maximum is NOT a raise
    cases.append(ramp.reshape(30, 100))
    for m in cases:
        assert np.array_equal(ctx.write_mat(m), oracle.write_mat(m)), m.shape
    d = torch.from_numpy(cases[3]).cuda()
    out = torch.empty(d.shape, dtype=torch.uint8, device="cuda")
    ctx.set_stream(torch.cuda.current_stream())
    ctx.fl_to_ch2_dev(d, out, -15, 0, d.numel())
    c = np.float32(160.0) * (cases[3] - np.float32(-15)) / np.float32(15)
    # (unsigned char) of a float is cvt.rzi on the device: truncation, negatives (the occlusion sentinel) saturate to 0
    want = np.where(c > 255, 255, np.clip(np.trunc(c), 0, 255)).astype(np.uint8)
    assert np.array_equal(out.cpu().numpy(), want)


# ---- fused pipeline ---------------------------------------------------------------------------
def test_fused_pipeline_tsukuba(ctx, oracle, tsukuba):
    L, R, gl, gr = tsukuba
    out = ctx.pipeline(L, R)
    assert np.array_equal(out["gray_left"], gl) and np.array_equal(out["gray_right"], gr)
    ref = oracle.pipeline_gray(gl, gr, -15, 16, oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads()),
                               want_second=True)
    check_fused_vs_oracle(out, ref, margin_tau=2e-4, min_agree=0.9999)
    # exact integer guide statistics: the mean image equals the exact-mode oracle's
    assert np.array_equal(out["mean_left"], ref["meanL"])
    # occlusion / fill are exact functions of the labels the kernel produced
    occ = oracle.detect_occlusion(out["disp_left"], out["disp_right"], -115)
    assert np.array_equal(out["occlusion"], occ)
    assert np.array_equal(out["filled"], oracle.fill_occlusion(occ, -15))
    # against the reference's own golden label maps (float32-SAT noise: SURVEY H1 expects ~99.995 %)
    gold_l = O.load_png("disparity_mapl.png").astype(int) // 17 - 15
    gold_r = O.load_png("disparity_mapr.png").astype(int) // 17
    assert (out["disp_left"].astype(int) == gold_l).mean() > 0.9999
    assert (out["disp_right"].astype(int) == gold_r).mean() > 0.9999
    gold_fill = O.load_png("occlu_mapl_filled.png").astype(int) // 17 - 15
    assert (out["filled"].astype(int) == gold_fill).mean() > 0.999


@pytest.mark.parametrize("w,h,dmin,dmax", [
    (230, 90, -20, 0),     # two strips, one band
    (216, 64, -3, 0),      # exactly one strip
    (217, 70, -5, 2),      # one pixel into a second strip, positive and negative d
    (64, 40, -7, 0),       # narrower than a strip
    (19, 19, -2, 0),       # smaller than the window in both directions
    (500, 150, -33, 0),    # D not a multiple of 4 (34 slices), several chunks
    (300, 30, 2, 9),       # dmin > 0
    (97, 11, 0, 0),        # D = 1
])
def test_fused_pipeline_shapes(ctx, oracle, w, h, dmin, dmax):
    size_d = dmax - dmin + 1
    L, R = synth.make_pair(w, h, max(size_d, 2), seed=w + h)
    p = api.default_params(dmin=dmin, dmax=dmax)
    out = ctx.pipeline(L, R, p)
    ref = oracle.pipeline_gray(L, R, dmin, size_d, oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads()),
                               want_second=True)
    check_fused_vs_oracle(out, ref, margin_tau=2e-4, min_agree=0.999)
    occ = oracle.detect_occlusion(out["disp_left"], out["disp_right"], dmin - 100)
    assert np.array_equal(out["occlusion"], occ)
    assert np.array_equal(out["filled"], oracle.fill_occlusion(occ, dmin))


def test_fused_tie_break_constant_image(ctx, oracle):
    """Constant pair: every slice whose whole (cascaded) window is in range filters to q == 0
    exactly, so the label must be the LAST such slice (guidedFilter.cu:406 `>=`): d = 0 everywhere
    for the left view, and the oracle's tie resolution pixel for pixel for the right view."""
    img = np.full((60, 250), 91, np.uint8)
    out = ctx.pipeline(img, img, api.default_params(dmin=-11, dmax=0))
    assert np.all(out["disp_left"] == 0)
    ref = oracle.pipeline_gray(img, img, -11, 12, oracle.params(box_mode=O.BOX_EXACT))
    assert np.array_equal(out["disp_right"], ref["dR"])
    assert np.all(out["disp_right"][:, :200] == 11) and np.all(out["disp_right"][:, -1] == 0)
    assert np.array_equal(out["filled"], ref["filled"])


def test_fused_synthetic_staircase_recovers_disparity(ctx):
    size_d = 64
    L, R = synth.make_pair(640, 256, size_d, seed=5)
    out = ctx.pipeline(L, R, api.default_params(dmin=-(size_d - 1), dmax=0))
    truth = -synth.delta_rows(256, size_d)[:, None].astype(np.float32)
    inner = out["disp_left"][:, size_d + 10:-10]
    assert (inner == np.broadcast_to(truth, out["disp_left"].shape)[:, size_d + 10:-10]).mean() > 0.9


def test_view_disparity_dev_and_lr_fill_dev(ctx, oracle, tsukuba):
    torch = pytest.importorskip("torch")
    _, _, gl, gr = tsukuba
    h, w = gl.shape
    dgl, dgr = torch.from_numpy(gl).cuda(), torch.from_numpy(gr).cuda()
    best = torch.empty((h, w), dtype=torch.float32, device="cuda")
    disp = torch.empty_like(best)
    ctx.set_stream(torch.cuda.current_stream())
    ctx.view_disparity_dev(dgl, dgr, w, h, -15, 16, best, disp)
    dispr = torch.empty_like(best)
    ctx.view_disparity_dev(dgr, dgl, w, h, 0, 16, None, dispr)
    occ, filled = torch.empty_like(best), torch.empty_like(best)
    ctx.lr_check_fill_dev(disp, dispr, w, h, -115, -15.0, occ, filled)
    torch.cuda.synchronize()
    out = ctx.pipeline(gl, gr)
    assert np.array_equal(disp.cpu().numpy(), out["disp_left"])
    assert np.array_equal(dispr.cpu().numpy(), out["disp_right"])
    assert np.array_equal(best.cpu().numpy(), out["best_left"])
    assert np.array_equal(occ.cpu().numpy(), out["occlusion"]) and np.array_equal(filled.cpu().numpy(), out["filled"])


def test_strip_mode_matches_whole_frame(ctx, oracle):
    """Row strips with 2*radius halo rows: labels of the strips, concatenated, equal
    the whole-frame labels; first-stage sums are exact so only second-stage float order differs."""
    torch = pytest.importorskip("torch")
    w, h, size_d = 300, 200, 24
    L, R = synth.make_pair(w, h, size_d, seed=11)
    p = api.default_params(dmin=-(size_d - 1), dmax=0)
    whole = ctx.pipeline(L, R, p)
    halo = ctx.strip_halo_rows(p)
    assert halo == 18
    ctx.set_stream(torch.cuda.current_stream())
    parts = {k: [] for k in ("disp_left", "disp_right", "filled", "best_left")}
    bounds = [0, 64, 128, 200]
    for y0, y1 in zip(bounds[:-1], bounds[1:]):
        top, bot = min(halo, y0), min(halo, h - y1)
        dl = torch.from_numpy(L[y0 - top:y1 + bot].copy()).cuda()
        dr = torch.from_numpy(R[y0 - top:y1 + bot].copy()).cuda()
        outs = {k: torch.empty((y1 - y0, w), dtype=torch.float32, device="cuda") for k in parts}
        ctx.pipeline_strip_dev(dl, dr, 1, w, dict(y0=y0, rows=y1 - y0, halo_top=top, halo_bot=bot, frame_h=h), outs, p)
        torch.cuda.synchronize()
        for k in parts:
            parts[k].append(outs[k].cpu().numpy())
    for k in ("disp_left", "disp_right", "filled"):
        got = np.concatenate(parts[k], 0)
        assert (got == whole[k]).mean() > 0.9999, k
    assert np.allclose(np.concatenate(parts["best_left"], 0), whole["best_left"], rtol=1e-5, atol=1e-6)


def test_unsupported_requests_fail_loudly(ctx):
    torch = pytest.importorskip("torch")
    img = np.zeros((40, 40), np.uint8)
    with pytest.raises(S.StereoB200Error):
        ctx.pipeline(img, img, api.default_params(dmin=3, dmax=1))
    # the fused-kernel-only entry point refuses what the fused kernel is not built for
    d = torch.zeros((40, 40), dtype=torch.uint8, device="cuda")
    out = torch.empty((40, 40), dtype=torch.float32, device="cuda")
    with pytest.raises(S.StereoB200Error, match="radius"):
        ctx.view_disparity_dev(d, d, 40, 40, -3, 4, out, out, params=api.default_params(radius=4))
    with pytest.raises(S.StereoB200Error, match="lattice"):
        ctx.view_disparity_dev(d, d, 40, 40, -3, 4, out, out, params=api.default_params(alpha=0.8731))


@pytest.mark.parametrize("kw", [dict(radius=4), dict(radius=12), dict(alpha=0.8731), dict(th_color=5.5), dict(eps=25.0)])
def test_pipeline_other_parameters_run_staged_or_fused(ctx, oracle, kw):
    """Parameters the fused kernel is not built for (another radius, a cost without an exact integer lattice)
    still run through sb200_pipeline -- on the GPU, through the stage kernels -- and match the oracle."""
    w, h, size_d = 150, 80, 9
    L, R = synth.make_pair(w, h, size_d, seed=77)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, **kw)
    # the caller can see which path a parameter set takes: the reference's macros run on the tensor-core kernel
    assert ctx.pipeline_path(api.default_params()) == 2
    assert ctx.pipeline_path(p) == (0 if ("radius" in kw or "alpha" in kw) else ctx.pipeline_path(p))
    assert ctx.pipeline_path(p) in (0, 1, 2)
    out = ctx.pipeline(L, R, p)
    okw = dict(kw)
    ref = oracle.pipeline_gray(L, R, -(size_d - 1), size_d,
                               oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads(), **okw), want_second=True)
    check_fused_vs_oracle(out, ref, margin_tau=2e-4, min_agree=0.999)
    occ = oracle.detect_occlusion(out["disp_left"], out["disp_right"], -(size_d - 1) - 100)
    assert np.array_equal(out["filled"], oracle.fill_occlusion(occ, -(size_d - 1)))


def test_compat_header_program_runs(ctx):
    """include/stereo_b200_compat.hpp: main.cu's call sequence with the reference's C++ signatures, linked
    against libstereo_b200.so, compiled and run on this box."""
    import shutil
    import subprocess

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    cxx = shutil.which("g++") or "/usr/bin/g++"
    exe = os.path.join(root, "gpurun_out", "compat_smoke")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.join(root, "stereo_matching_cuda_b200")
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "compat_smoke.cpp"),
                    "-o", exe, "-L", libdir, "-lstereo_b200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "labelled -4" in r.stdout


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libref.so not built")
def test_against_reference_gpu_code(ctx, oracle, tsukuba):
    """The reference's own host functions + kernels recompiled for sm_100a (oracle/_ref), run in a
    child process under a timeout (its rowSum/colSum call __syncthreads after a divergent return)."""
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "run_ref_gpu.py")], capture_output=True, text=True, timeout=240)
    if r.returncode != 0:
        pytest.skip("reference GPU code did not run here: " + (r.stderr or r.stdout)[-300:])
    ref = dict(np.load(os.path.join(here, "..", "gpurun_out", "ref_gpu_tsukuba.npz")))
    _, _, gl, gr = tsukuba
    assert np.array_equal(ref["gray_l"], gl)
    assert np.array_equal(ref["cost_l"], ctx.compute_cost(gl, gr, -15))
    p = api.default_params(box_mode=S.BOX_SAT)
    best = np.full(gl.shape, oracle.best_init(), np.float32)
    dmap = np.zeros(gl.shape, np.float32)
    ctx.compute_guided_filter(gl, ref["cost_l"], best, dmap, -15, p)
    assert (dmap == ref["dmap_l"]).mean() > 0.9999  # its device code is fma-contracted, ours is not
    assert np.allclose(best, ref["best_l"], rtol=1e-3)  # contraction of mIp - mean*mp cancels ~3 digits
    out = ctx.pipeline(gl, gr)
    assert (out["disp_left"] == ref["dmap_l"]).mean() > 0.9999
    assert (out["filled"] == ref["filled"]).mean() > 0.999


def test_fused_full_size_1080p_d256(ctx, oracle):
    """BASELINE.json's headline shape: 1920x1080, D=256.  BOTH views against the exact-mode oracle (all host
    threads, ~15 s per view): labels, and the best cost within 1e-4 relative with a floor of 1e-2 (the median best
    cost of this pair is 0.02), plus size-independent properties on the whole pipeline output."""
    torch = pytest.importorskip("torch")
    w, h, size_d = 1920, 1080, 256
    L, R = synth.make_pair(w, h, size_d, seed=0)
    p = api.default_params(dmin=-(size_d - 1), dmax=0)
    assert ctx.gray_kernel == 1  # the tensor-core kernel is the default
    out = ctx.pipeline(L, R, p, want=("disp_left", "disp_right", "best_left", "best_right", "occlusion", "filled"))
    po = oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads())
    record = {}
    for view, (guide, other, dmin, kd, kb) in {"left": (L, R, -(size_d - 1), "disp_left", "best_left"),
                                               "right": (R, L, 0, "disp_right", "best_right")}.items():
        best, dmap, _, second = oracle.view_disparity(guide, other, size_d, dmin, po, want_second=True)
        same = out[kd] == dmap
        err = np.abs(out[kb] - best)
        rel = err / np.maximum(np.abs(best), 1e-2)
        record[view] = {"label_agreement": float(same.mean()), "max_abs_err": float(err.max()),
                        "p999_abs_err": float(np.quantile(err, 0.999)), "max_rel_err_floor1e-2": float(rel.max()),
                        "max_rel_err_floor0.1": float((err / np.maximum(np.abs(best), 0.1)).max()),
                        "median_best": float(np.median(best)), "min_best": float(best.min())}
        assert same.mean() >= 0.999, (view, same.mean())
        assert np.all(same[(second - best) > 2e-4]), view
        assert rel.max() < RTOL_BEST, (view, rel.max(), err.max())
    import json
    os.makedirs(os.path.join(os.path.dirname(__file__), "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "fullsize_parity.json"), "w") as f:
        json.dump(record, f)
    # properties: labels inside the search range, occlusion map is labels-or-sentinel, fill leaves
    # no sentinel, is idempotent and only changes occluded pixels
    assert out["disp_left"].min() >= -(size_d - 1) and out["disp_left"].max() <= 0
    assert out["disp_right"].min() >= 0 and out["disp_right"].max() <= size_d - 1
    occ = out["occlusion"]
    sent = -(size_d - 1) - 100
    assert np.array_equal(occ[occ != sent], out["disp_left"][occ != sent])
    assert np.all(out["filled"] >= -(size_d - 1))
    assert np.array_equal(out["filled"][occ != sent], occ[occ != sent])
    again = out["filled"].copy()
    ctx.fill_occlusion(again, -(size_d - 1))
    assert np.array_equal(again, out["filled"])
    assert np.array_equal(occ, oracle.detect_occlusion(out["disp_left"], out["disp_right"], sent))
    # the staircase ground truth is recovered away from the band edges and image borders
    truth = -synth.delta_rows(h, size_d)[:, None].astype(np.float32)
    core = np.zeros((h, w), bool)
    core[:, size_d + 20:-20] = True
    assert (out["disp_left"][core] == np.broadcast_to(truth, (h, w))[core]).mean() > 0.9


def test_shuffle_kernel_matches_oracle(oracle):
    """k_fused_cvf (warp-shuffle window sums, round 1's kernel) stays selectable (SB200_GRAY_KERNEL=shfl or
    Context.gray_kernel = 0) and in parity: same checks as the default kernel on strips, bands, chunks and a partial group."""
    with S.Context(0) as c:
        c.gray_kernel = 0
        assert c.gray_kernel == 0
        for (w, h, dmin, dmax) in ((230, 90, -20, 0), (500, 150, -33, 0), (300, 30, 2, 9), (19, 19, -2, 0)):
            size_d = dmax - dmin + 1
            L, R = synth.make_pair(w, h, max(size_d, 2), seed=w + h)
            out = c.pipeline(L, R, api.default_params(dmin=dmin, dmax=dmax))
            ref = oracle.pipeline_gray(L, R, dmin, size_d, oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads()),
                                       want_second=True)
            check_fused_vs_oracle(out, ref, margin_tau=2e-4, min_agree=0.999)
        with pytest.raises(S.StereoB200Error):
            c.gray_kernel = 7


def test_two_kernels_agree_at_1080p(ctx):
    """The two gray-guide kernels are independent implementations (different tilings, different box-sum machinery):
    at 1920x1080, D=256 their labels agree on > 99.99 % of the pixels and their best costs to 1e-5 absolute."""
    w, h, size_d = 1920, 1080, 256
    L, R = synth.make_pair(w, h, size_d, seed=0)
    p = api.default_params(dmin=-(size_d - 1), dmax=0)
    want = ("disp_left", "disp_right", "best_left", "best_right")
    a = ctx.pipeline(L, R, p, want=want)
    with S.Context(0) as c:
        c.gray_kernel = 0
        b = c.pipeline(L, R, p, want=want)
    for k in ("disp_left", "disp_right"):
        assert (a[k] == b[k]).mean() > 0.9999, k
    for k in ("best_left", "best_right"):
        assert np.abs(a[k] - b[k]).max() < 1e-5, k


def test_context_on_another_device_leaves_current_device_alone(oracle):
    """every entry point runs on the context's device and restores the caller's."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    L, R = synth.make_pair(200, 60, 8, seed=5)
    p = api.default_params(dmin=-7, dmax=0)
    with S.Context(0) as c0, S.Context(1) as c1:
        a = c0.pipeline(L, R, p)
        b = c1.pipeline(L, R, p)
        assert torch.cuda.current_device() == 0
        d1 = torch.from_numpy(b["disp_left"]).to("cuda:1")
        occ = torch.empty_like(d1)
        c1.lr_check_fill_dev(d1, torch.from_numpy(b["disp_right"]).to("cuda:1"), 200, 60, -107, -7.0, occ, None, p)
        c1.synchronize()
        assert torch.cuda.current_device() == 0
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(occ.cpu().numpy(), a["occlusion"])


def test_cli_reproduces_the_twelve_golden_pngs(tsukuba):
    """the driver's 12 output images decode to exactly the reference's checked-in PNGs."""
    from stereo_matching_cuda_b200 import cli

    L, R, _, _ = tsukuba
    files = cli.run(L, R)
    assert len(files) == 12
    for name, img in files.items():
        assert np.array_equal(img, O.load_png(name)), name
    fused = cli.run(L, R, fused=True)
    for name in ("image_left.png", "image_right.png", "cost_lminus15.png", "cost_rminus15.png"):
        assert np.array_equal(fused[name], O.load_png(name)), name
    for name in ("disparity_mapl.png", "disparity_mapr.png", "occlu_mapl_filled.png"):
        assert (fused[name] == O.load_png(name)).mean() > 0.999, name
    sub = cli.run(L, R, subpixel=True)  # beyond the reference: two more files, the twelve unchanged
    assert len(sub) == 14 and all(np.array_equal(sub[k], fused[k]) for k in fused)
    sp = sub["disparity_mapl_subpixel.npy"]
    assert sp.dtype == np.float32 and sp.shape == L.shape[:2] and sp.min() >= -15.5 and sp.max() <= 0.5
    assert sub["disparity_mapl_subpixel.png"].dtype == np.uint8


@pytest.mark.parametrize("w,h,dmin,dmax", [(2, 1, -1, 0), (5, 3, -2, 1), (40, 1, -3, 0), (3, 50, 0, 2), (433, 65, -9, 0)])
def test_fused_pipeline_tiny_and_ragged(ctx, oracle, w, h, dmin, dmax):
    rng = np.random.default_rng(w * 100 + h)
    L = rng.integers(0, 256, (h, w), dtype=np.uint8)
    R = rng.integers(0, 256, (h, w), dtype=np.uint8)
    size_d = dmax - dmin + 1
    out = ctx.pipeline(L, R, api.default_params(dmin=dmin, dmax=dmax))
    ref = oracle.pipeline_gray(L, R, dmin, size_d, oracle.params(box_mode=O.BOX_EXACT), want_second=True)
    check_fused_vs_oracle(out, ref, margin_tau=2e-4, min_agree=0.99)


def test_batch_dev_equals_single_pairs(ctx):
    torch = pytest.importorskip("torch")
    w, h, size_d, n = 260, 80, 10, 3
    pairs = [synth.make_pair(w, h, size_d, seed=40 + i) for i in range(n)]
    p = api.default_params(dmin=-(size_d - 1), dmax=0)
    dl = torch.from_numpy(np.stack([a for a, _ in pairs])).cuda()
    dr = torch.from_numpy(np.stack([b for _, b in pairs])).cuda()
    outs = {k: torch.empty((n, h, w), dtype=torch.float32, device="cuda") for k in ("disp_left", "filled", "best_right")}
    ctx.set_stream(torch.cuda.current_stream())
    ctx.pipeline_batch_dev(dl, dr, 1, w, h, n, outs, p)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(pairs):
        one = ctx.pipeline(a, b, p, want=("disp_left", "filled", "best_right"))
        for k in outs:
            assert np.array_equal(outs[k][i].cpu().numpy(), one[k]), (i, k)


def test_invalid_arguments_are_reported_not_fatal(ctx):
    img = np.zeros((30, 30), np.uint8)
    for bad in (dict(radius=0), dict(box_mode=7)):
        with pytest.raises(S.StereoB200Error):
            ctx.pipeline(img, img, api.default_params(**bad))
    with pytest.raises(S.StereoB200Error, match="same size"):
        ctx.compute_cost(img, np.zeros((30, 31), np.uint8), -3)
    with pytest.raises(S.StereoB200Error):
        ctx.pipeline(np.zeros((30, 30, 2), np.uint8), np.zeros((30, 30, 2), np.uint8))
    # the context is still usable afterwards
    assert ctx.pipeline(img, img)["disp_left"].shape == (30, 30)


def test_full_size_c2_bike_shape_rgb(ctx, oracle):
    """BASELINE configs[1]: 3052x1968 colour input, D=16, full pipeline incl. fill, against the exact oracle."""
    w, h, size_d = 3052, 1968, 16
    L, R = synth.make_pair(w, h, size_d, channels=3, seed=2)
    out = ctx.pipeline(L, R, want=("gray_left", "gray_right", "disp_left", "disp_right", "best_left", "best_right",
                                   "occlusion", "filled"))
    gl, gr = oracle.rgb_to_gray(L), oracle.rgb_to_gray(R)
    assert np.array_equal(out["gray_left"], gl) and np.array_equal(out["gray_right"], gr)
    ref = oracle.pipeline_gray(gl, gr, -15, size_d, oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads()),
                               want_second=True)
    for lab, best, second, o_lab, o_best in (("dL", "bestL", "secondL", "disp_left", "best_left"),
                                             ("dR", "bestR", "secondR", "disp_right", "best_right")):
        same = out[o_lab] == ref[lab]
        assert same.mean() >= 0.999
        assert np.all(same[(ref[second] - ref[best]) > 1e-3])
        assert (np.abs(out[o_best] - ref[best]) / np.maximum(np.abs(ref[best]), 0.1)).max() < RTOL_BEST
    occ = oracle.detect_occlusion(out["disp_left"], out["disp_right"], -115)
    assert np.array_equal(out["occlusion"], occ)
    assert np.array_equal(out["filled"], oracle.fill_occlusion(occ, -15))


def test_full_size_c5_8k_properties(ctx, oracle):
    """BASELINE configs[4] shape: 7680x4320, D=512 on one GPU.  Too large for the CPU oracle's filter, so:
    size-independent properties, the exact occlusion/fill functions on the produced labels, and a
    sampled row band recomputed as a strip (with halos) that must reproduce the whole-frame labels."""
    torch = pytest.importorskip("torch")
    w, h, size_d = 7680, 4320, 512
    L, R = synth.make_pair(w, h, size_d, seed=0)
    p = api.default_params(dmin=-(size_d - 1), dmax=0)
    out = ctx.pipeline(L, R, p, want=("disp_left", "disp_right", "occlusion", "filled"))
    assert out["disp_left"].min() >= -(size_d - 1) and out["disp_left"].max() <= 0
    assert out["disp_right"].min() >= 0 and out["disp_right"].max() <= size_d - 1
    sent = -(size_d - 1) - 100
    occ = oracle.detect_occlusion(out["disp_left"], out["disp_right"], sent)
    assert np.array_equal(out["occlusion"], occ)
    assert np.array_equal(out["filled"], oracle.fill_occlusion(occ, -(size_d - 1)))
    truth = -synth.delta_rows(h, size_d)[:, None].astype(np.float32)
    core = np.zeros((h, w), bool)
    core[:, size_d + 20:-20] = True
    assert (out["disp_left"][core] == np.broadcast_to(truth, (h, w))[core]).mean() > 0.9
    # one 128-row strip in the middle of the frame, recomputed from its rows + 18-row halos
    y0, rows, halo = 2048, 128, 18
    dl = torch.from_numpy(L[y0 - halo:y0 + rows + halo].copy()).cuda()
    dr = torch.from_numpy(R[y0 - halo:y0 + rows + halo].copy()).cuda()
    so = {k: torch.empty((rows, w), dtype=torch.float32, device="cuda") for k in ("disp_left", "disp_right")}
    ctx.set_stream(torch.cuda.current_stream())
    ctx.pipeline_strip_dev(dl, dr, 1, w, dict(y0=y0, rows=rows, halo_top=halo, halo_bot=halo, frame_h=h), so, p)
    torch.cuda.synchronize()
    for k in so:
        assert (so[k].cpu().numpy() == out[k][y0:y0 + rows]).mean() > 0.9999, k
    # the same band against the EXACT oracle: rows y0-18 .. y0+rows+18 of both images are everything the
    # cascaded radius-9 filters of rows y0 .. y0+rows read (the band lies far from the frame's top and bottom, so every
    # window inside it is a full 19x19 window, as in the frame); the oracle's own border rows are discarded
    yb, ye = y0 - halo, y0 + rows + halo
    want = ("disp_left", "disp_right", "best_left", "best_right")
    band = ctx.pipeline(L[yb:ye], R[yb:ye], p, want=want)  # (a cheaper way to get best costs of these rows than the 8K frame)
    po = oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads())
    for guide, other, dmin, kd, kb in ((L, R, -(size_d - 1), "disp_left", "best_left"), (R, L, 0, "disp_right", "best_right")):
        best, dmap, _, second = oracle.view_disparity(guide[yb:ye], other[yb:ye], size_d, dmin, po, want_second=True)
        sl = slice(halo, halo + rows)
        same = out[kd][y0:y0 + rows] == dmap[sl]
        assert same.mean() >= 0.999, (kd, same.mean())
        assert np.all(same[(second[sl] - best[sl]) > 2e-4]), kd
        # the whole frame's labels; the best costs of the same rows come from the band call (same kernel, same rows,
        # interior windows), which the strip check above ties to the whole-frame run
        assert (band[kd][sl] == out[kd][y0:y0 + rows]).mean() > 0.9999
        rel = np.abs(band[kb][sl] - best[sl]) / np.maximum(np.abs(best[sl]), 1e-2)
        assert rel.max() < RTOL_BEST, (kb, rel.max())


def test_width_one_is_rejected(ctx):
    img = np.zeros((8, 1), np.uint8)
    with pytest.raises(S.StereoB200Error):
        ctx.pipeline(img, img)  # the reference's own gradient kernel reads out of bounds at w == 1


def test_fused_vs_staged_random_shapes(ctx):
    """Two independent CUDA implementations of the same path (fused kernel vs stage-by-stage kernels with
    double-accumulated sliding boxes) on random shapes and disparity ranges: exercises strips, bands,
    chunks and partial disparity groups without needing the CPU."""
    rng = np.random.default_rng(2026)
    for _ in range(10):
        w = int(rng.integers(24, 900))
        h = int(rng.integers(8, 260))
        size_d = int(rng.integers(1, 40))
        dmax = int(rng.integers(0, 3))
        dmin = dmax - size_d + 1
        L, R = synth.make_pair(w, h, max(size_d, 2), seed=int(rng.integers(1 << 20)))
        p = api.default_params(dmin=dmin, dmax=dmax)
        out = ctx.pipeline(L, R, p, want=("disp_left", "best_left", "disp_right", "best_right"))
        for guide, other, dm, kd, kb in ((L, R, dmin, "disp_left", "best_left"), (R, L, -dmax, "disp_right", "best_right")):
            cost = ctx.compute_cost(guide, other, dm, p)
            best = np.full((h, w), 3.3961514e38, np.float32)
            dmap = np.zeros((h, w), np.float32)
            ctx.compute_guided_filter(guide, cost, best, dmap, dm, p)  # box_mode SLIDING: double accumulation
            assert (np.abs(out[kb] - best) / np.maximum(np.abs(best), 0.1)).max() < RTOL_BEST, (w, h, dmin, dmax)
            assert (out[kd] == dmap).mean() > 0.998, (w, h, dmin, dmax, (out[kd] == dmap).mean())


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,size_d,radius", [(384, 288, 16, 19), (150, 70, 12, 5), (41, 23, 3, 19), (300, 64, 100, 7)])
def test_weighted_median_equals_the_oracle(oracle, w, h, size_d, radius):
    """SURVEY 8f.3 (beyond the reference): the weighted median of the marked pixels is defined with integer weights
    (stereo_b200.h), so the CUDA path must equal the oracle's restatement bit for bit"""
    S = pytest.importorskip("stereo_matching_cuda_b200")
    from stereo_matching_cuda_b200 import api

    L, R = synth.make_pair(w, h, max(size_d, 2), seed=w + radius)
    dmin = -(size_d - 1)
    p = api.default_params(dmin=dmin, dmax=0)
    with S.Context(0) as ctx:
        out = ctx.pipeline(L, R, p, want=("occlusion", "filled", "gray_left"))
        got = ctx.weighted_median(L, out["occlusion"], out["filled"], p, radius=radius)
        with pytest.raises(S.StereoB200Error):
            ctx.weighted_median(L, out["occlusion"], out["filled"], p, radius=40)
    marked = out["occlusion"].astype(np.int32) < dmin
    assert marked.any()
    want = oracle.weighted_median(L, out["occlusion"], out["filled"], dmin, size_d, radius=radius)
    assert np.array_equal(got, want)
    assert np.array_equal(got[~marked], out["filled"][~marked])
    assert got.min() >= dmin and got.max() <= 0


def _oracle_filtered_volume(oracle, guide, other, size_d, dmin):
    """the filtered cost q of every slice (compute_q, guidedFilter.cu:363-369), exact box mode: what the reference's
    per-slice loop (guidedFilter.cu:171-238) holds before dispSelectOnGPU consumes it"""
    po = oracle.params(box_mode=O.BOX_EXACT, nthreads=1)
    I, mI, vI, _ = oracle.guide_stats(guide, po)
    cost = oracle.cost_volume(guide, other, size_d, dmin, po)
    return np.stack([oracle.guided_slice(I, mI, vI, cost[k], po)[0] for k in range(size_d)])


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,size_d", [(300, 120, 24), (111, 45, 9), (257, 33, 67)])
def test_filtered_volume_and_subpixel_refinement(oracle, w, h, size_d):
    """SURVEY 8f.3 (beyond the reference).  (1) The volume the fused kernel keeps on request equals the oracle's filtered
    slices (1e-4 relative, 1e-2 floor) and its per-pixel minimum IS the best-cost map, bit for bit.  (2) The refinement
    kernel equals the oracle's restatement bit for bit on that volume.  (3) Through the pipeline: marked pixels keep
    their filled label, the others move by at most half a label."""
    import torch

    L, R = synth.make_pair(w, h, max(size_d, 2), seed=3 * w + size_d)
    dmin = -(size_d - 1)
    p = api.default_params(dmin=dmin, dmax=0)
    dev = torch.device("cuda:0")
    with S.Context(0) as ctx:
        dl, dr = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
        vol = torch.full((size_d, h, w), float("nan"), dtype=torch.float32, device=dev)
        best, disp = (torch.empty((h, w), dtype=torch.float32, device=dev) for _ in range(2))
        ctx.view_volume_dev(dl, dr, w, h, dmin, size_d, vol, best, disp, params=p)
        ref_best, ref_disp = (torch.empty((h, w), dtype=torch.float32, device=dev) for _ in range(2))
        ctx.view_disparity_dev(dl, dr, w, h, dmin, size_d, ref_best, ref_disp, params=p)
        sub = torch.empty((h, w), dtype=torch.float32, device=dev)
        ctx.subpixel_refine_dev(vol, disp, sub, w, h, dmin, size_d)
        ctx.synchronize()
        vol_h, best_h, disp_h, sub_h = vol.cpu().numpy(), best.cpu().numpy(), disp.cpu().numpy(), sub.cpu().numpy()
        # keeping the volume does not change the kernel's other outputs
        assert np.array_equal(best_h, ref_best.cpu().numpy()) and np.array_equal(disp_h, ref_disp.cpu().numpy())
        out = ctx.pipeline(L, R, p, want=("disp_left", "occlusion", "filled", "subpixel_left"))
        plain = ctx.pipeline(L, R, p, want=("disp_left", "occlusion", "filled"))
    assert not np.isnan(vol_h).any()
    q = _oracle_filtered_volume(oracle, L, R, size_d, dmin)
    rel = np.abs(vol_h - q) / np.maximum(np.abs(q), 1e-2)
    assert rel.max() < RTOL_BEST, rel.max()
    assert np.array_equal(vol_h.min(axis=0), best_h)
    k = (disp_h - dmin).astype(np.int64)
    assert np.array_equal(np.take_along_axis(vol_h, k[None], 0)[0], best_h)
    # (2) bit-exact refinement on the same volume
    want = oracle.subpixel_refine(vol_h, disp_h, dmin)
    assert np.array_equal(sub_h.view(np.uint32), want.view(np.uint32))
    assert np.abs(sub_h - disp_h).max() <= 0.5
    assert (sub_h != disp_h).mean() > 0.3  # the synthetic pair has texture: most interior labels do move
    edge = (k == 0) | (k == size_d - 1)
    assert np.array_equal(sub_h[edge], disp_h[edge])
    # (3) pipeline output
    for name in ("disp_left", "occlusion", "filled"):
        assert np.array_equal(out[name], plain[name]), name
    marked = out["occlusion"].astype(np.int32) < dmin
    assert np.array_equal(out["subpixel_left"][marked], out["filled"][marked])
    assert np.array_equal(out["subpixel_left"], oracle.subpixel_refine(vol_h, out["disp_left"], dmin, out["occlusion"], out["filled"]))


@pytest.mark.gpu
def test_subpixel_rgb_guide_and_refusals(ctx):
    """the RGB tensor-core kernel keeps its volume the same way: on a grey-valued colour pair the RGB guide is the gray
    guide with eps/3 (the identity tests/test_rgb_guide.py pins the oracle's RGB port with), so the two refinements agree;
    paths that never hold the volume (staged parameters) refuse"""
    w, h, size_d = 280, 90, 20
    Lg, Rg = synth.make_pair(w, h, size_d, seed=77)
    L3, R3 = (np.repeat(a[..., None], 3, axis=2).copy() for a in (Lg, Rg))
    dmin = -(size_d - 1)
    want = ("disp_left", "occlusion", "filled", "subpixel_left")
    rgb = ctx.pipeline(L3, R3, api.default_params(dmin=dmin, dmax=0, guide_mode=S.GUIDE_RGB), want=want)
    gray = ctx.pipeline(Lg, Rg, api.default_params(dmin=dmin, dmax=0, eps=6.5025 / 3), want=want)
    marked = rgb["occlusion"].astype(np.int32) < dmin
    assert np.array_equal(rgb["subpixel_left"][marked], rgb["filled"][marked])
    assert np.abs(rgb["subpixel_left"] - rgb["disp_left"])[~marked].max() <= 0.5
    kept = ~marked & ~(gray["occlusion"].astype(np.int32) < dmin)
    assert kept.mean() > 0.2
    same = (rgb["disp_left"] == gray["disp_left"]) & kept
    assert same.sum() > 0.99 * kept.sum()
    diff = np.abs(rgb["subpixel_left"] - gray["subpixel_left"])[same]
    # the vertex amplifies the 1e-4 differences between the two filters where the three costs are nearly collinear
    assert np.median(diff) < 5e-3 and np.quantile(diff, 0.99) < 5e-2, (np.median(diff), np.quantile(diff, 0.99))
    assert (rgb["subpixel_left"] != rgb["disp_left"]).mean() > 0.3
    p2 = api.default_params(dmin=-5, dmax=0, radius=4)  # staged path
    with pytest.raises(S.StereoB200Error):
        ctx.pipeline(Lg, Rg, p2, want=("disp_left", "subpixel_left"))


@pytest.mark.gpu
def test_batch_int16_labels_equal_the_float_maps(ctx):
    """sb200_pipeline_batch_i16: the four label maps as int16 are the float maps, value for value; other
    outputs may ride along as floats; a map asked for in both forms and labels outside int16 are refused"""
    n, w, h, size_d = 3, 333, 77, 40
    pairs = [synth.make_pair(w, h, size_d, seed=50 + i) for i in range(n)]
    Ls, Rs = np.stack([a for a, _ in pairs]), np.stack([b for _, b in pairs])
    p = api.default_params(dmin=-(size_d - 1), dmax=0)
    want = ("disp_left", "disp_right", "occlusion", "filled", "best_left")
    f = ctx.pipeline_batch(Ls, Rs, p, want=want)
    i16 = ctx.pipeline_batch(Ls, Rs, p, want=want, labels_i16=True)
    for k in ("disp_left", "disp_right", "occlusion", "filled"):
        assert i16[k].dtype == np.int16
        assert np.array_equal(i16[k].astype(np.float32), f[k]), k
    assert i16["best_left"].dtype == np.float32 and np.array_equal(i16["best_left"], f["best_left"])
    assert (f["occlusion"] == -(size_d - 1) - 100).any()
    with pytest.raises(S.StereoB200Error):
        ctx.pipeline_batch(Ls, Rs, api.default_params(dmin=-40000, dmax=-39990), want=("filled",), labels_i16=True)


@pytest.mark.gpu
@pytest.mark.parametrize("rgb", [False, True])
def test_fused_outputs_are_deterministic_over_repeated_runs(ctx, rgb):
    """The fused kernels hand rows between five warp roles through mbarriers, Tensor Memory and shared-memory rings; a missing
    ordering would show as run-to-run differences long before it shows as a wrong pixel.  Twelve runs of one pair (several
    disparity groups, bands and strips) must agree in every bit of every output."""
    w, h, size_d = 700, 400, 72
    L, R = synth.make_pair(w, h, size_d, channels=3 if rgb else 1, seed=11)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB if rgb else S.GUIDE_GRAY)
    want = ("disp_left", "disp_right", "best_left", "best_right", "occlusion", "filled")
    first = ctx.pipeline(L, R, p, want=want)
    for _ in range(11):
        again = ctx.pipeline(L, R, p, want=want)
        for k in want:
            assert np.array_equal(first[k].view(np.uint32), again[k].view(np.uint32)), k


@pytest.mark.gpu
def test_subpixel_in_strip_mode_equals_the_whole_frame(ctx):
    """the kept volume follows the strip geometry (rows_out x w planes): a strip's sub-pixel map equals the whole frame's rows
    wherever the integer labels agree (strips differ from the whole frame only by second-stage float summation order)"""
    torch = pytest.importorskip("torch")
    w, h, size_d = 300, 200, 20
    L, R = synth.make_pair(w, h, size_d, seed=9)
    dmin = -(size_d - 1)
    p = api.default_params(dmin=dmin, dmax=0)
    whole = ctx.pipeline(L, R, p, want=("disp_left", "occlusion", "filled", "subpixel_left"))
    dev = torch.device("cuda:0")
    halo = ctx.strip_halo_rows(p)
    y0, rows = 70, 64
    top, bot = min(halo, y0), min(halo, h - (y0 + rows))
    dl = torch.from_numpy(np.ascontiguousarray(L[y0 - top:y0 + rows + bot])).to(dev)
    dr = torch.from_numpy(np.ascontiguousarray(R[y0 - top:y0 + rows + bot])).to(dev)
    names = ("disp_left", "disp_right", "occlusion", "filled", "subpixel_left")
    outs = {k: torch.empty((rows, w), dtype=torch.float32, device=dev) for k in names}
    ctx.pipeline_strip_dev(dl, dr, 1, w, dict(y0=y0, rows=rows, halo_top=top, halo_bot=bot, frame_h=h), outs, p)
    ctx.synchronize()
    got = {k: v.cpu().numpy() for k, v in outs.items()}
    same = got["disp_left"] == whole["disp_left"][y0:y0 + rows]
    assert same.mean() > 0.999
    assert np.array_equal(got["occlusion"][same], whole["occlusion"][y0:y0 + rows][same])
    d = np.abs(got["subpixel_left"] - whole["subpixel_left"][y0:y0 + rows])[same & (got["occlusion"] >= dmin)]
    assert np.median(d) < 1e-4 and np.quantile(d, 0.99) < 1e-2, (np.median(d), np.quantile(d, 0.99))
    assert (got["subpixel_left"] != got["disp_left"]).mean() > 0.3
