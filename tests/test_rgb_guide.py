"""RGB guide.  The reference has no colour-guide path, so parity is unpinned by
construction: the oracle is checked through the identity "RGB path on a grey-valued RGB image ==
gray path with eps/3" and the CUDA path is checked against that oracle."""
import numpy as np
import pytest

import _oracle as O
import synth


def test_oracle_rgb_equals_gray_with_eps_over_3(oracle):
    L, R = synth.make_pair(90, 60, 8, seed=21)
    rgb = np.repeat(L[..., None], 3, axis=2).copy()
    size_d, dmin = 8, -7
    p_rgb = oracle.params(box_mode=O.BOX_EXACT, nthreads=4)
    b_rgb, d_rgb, _ = oracle.view_disparity_rgb(rgb, L, R, size_d, dmin, p_rgb)
    p_gray = oracle.params(box_mode=O.BOX_EXACT, nthreads=4, eps=6.5025 / 3)
    b_gray, d_gray, _, s_gray = oracle.view_disparity(L, R, size_d, dmin, p_gray, want_second=True)
    assert np.allclose(b_rgb, b_gray, rtol=2e-4, atol=1e-5)
    margin = s_gray - b_gray
    assert np.array_equal(d_rgb[margin > 1e-3], d_gray[margin > 1e-3])
    assert (d_rgb == d_gray).mean() > 0.995


@pytest.mark.gpu
def test_gpu_rgb_guide_matches_oracle(oracle):
    S = pytest.importorskip("stereo_matching_cuda_b200")
    from stereo_matching_cuda_b200 import api

    w, h, size_d = 150, 70, 12
    L, R = synth.make_pair(w, h, size_d, channels=3, seed=5)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
    with S.Context(0) as ctx:
        out = ctx.pipeline(L, R, p)
        with pytest.raises(S.StereoB200Error):
            ctx.pipeline(L[..., 0].copy(), R[..., 0].copy(), p)  # an RGB guide needs colour input
    gl, gr = oracle.rgb_to_gray(L), oracle.rgb_to_gray(R)
    assert np.array_equal(out["gray_left"], gl)
    po = oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads())
    bl, dl, sl = oracle.view_disparity_rgb(L, gl, gr, size_d, -(size_d - 1), po, want_second=True)
    br, dr, sr = oracle.view_disparity_rgb(R, gr, gl, size_d, 0, po, want_second=True)
    for got_d, got_b, d, b, s in ((out["disp_left"], out["best_left"], dl, bl, sl),
                                  (out["disp_right"], out["best_right"], dr, br, sr)):
        rel = np.abs(got_b - b) / np.maximum(np.abs(b), 1e-2)
        assert rel.max() < 1e-4
        assert np.array_equal(got_d[(s - b) > 2e-4], d[(s - b) > 2e-4])
        assert (got_d == d).mean() > 0.999
    occ = oracle.detect_occlusion(out["disp_left"], out["disp_right"], -(size_d - 1) - 100)
    assert np.array_equal(out["occlusion"], occ)
    assert np.array_equal(out["filled"], oracle.fill_occlusion(occ, -(size_d - 1)))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,size_d", [(470, 130, 21), (216, 40, 4), (33, 25, 3)])
def test_gpu_rgb_fused_shapes_and_staged_agree(oracle, w, h, size_d):
    """fused RGB kernel (default) vs the oracle on several tilings, and vs the staged RGB path"""
    S = pytest.importorskip("stereo_matching_cuda_b200")
    from stereo_matching_cuda_b200 import api

    L, R = synth.make_pair(w, h, max(size_d, 2), channels=3, seed=w)
    dmin = -(size_d - 1)
    p = api.default_params(dmin=dmin, dmax=0, guide_mode=S.GUIDE_RGB)
    with S.Context(0) as ctx:
        out = ctx.pipeline(L, R, p, want=("disp_left", "disp_right", "best_left", "best_right", "filled"))
    gl, gr = oracle.rgb_to_gray(L), oracle.rgb_to_gray(R)
    po = oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads())
    for guide, g1, g2, dm, kd, kb in ((L, gl, gr, dmin, "disp_left", "best_left"), (R, gr, gl, 0, "disp_right", "best_right")):
        b, d, s = oracle.view_disparity_rgb(guide, g1, g2, size_d, dm, po, want_second=True)
        assert (np.abs(out[kb] - b) / np.maximum(np.abs(b), 0.1)).max() < 1e-4
        assert np.array_equal(out[kd][(s - b) > 2e-4], d[(s - b) > 2e-4])
        assert (out[kd] == d).mean() > 0.999


def _rgb_pipeline_with_kernel(S, api, monkeypatch, which, L, R, p):
    monkeypatch.setenv("SB200_RGB_KERNEL", str(which))  # read by sb200_ctx_create
    with S.Context(0) as ctx:
        return ctx.pipeline(L, R, p, want=("disp_left", "disp_right", "best_left", "best_right", "occlusion", "filled"))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,size_d", [(470, 130, 21), (216, 40, 4), (33, 25, 3), (217, 19, 1), (900, 330, 70)])
def test_gpu_rgb_three_stage_kernel_equals_two_stage_bitwise(monkeypatch, w, h, size_d):
    """k_fused_cvf_rgb3 (three warp roles, TMA operand ring) and k_fused_cvf_rgb (two roles, L1 loads) perform the
    same float operations in the same order on the same exact first-stage sums: every output is bit-identical."""
    S = pytest.importorskip("stereo_matching_cuda_b200")
    from stereo_matching_cuda_b200 import api

    L, R = synth.make_pair(w, h, max(size_d, 2), channels=3, seed=w + 1)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
    a = _rgb_pipeline_with_kernel(S, api, monkeypatch, 2, L, R, p)
    b = _rgb_pipeline_with_kernel(S, api, monkeypatch, 3, L, R, p)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32) if a[k].dtype == np.float32 else a[k],
                              b[k].view(np.uint32) if b[k].dtype == np.float32 else b[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,size_d", [(470, 130, 21), (216, 40, 4), (33, 25, 3), (217, 19, 1), (900, 330, 70)])
def test_gpu_rgb_tensor_core_kernel_agrees_with_the_shuffle_kernel(monkeypatch, w, h, size_d):
    """k_fused_mma_rgb (the default, SB200_RGB_KERNEL=4) against k_fused_cvf_rgb3 (=3): the first-stage sums of both are
    exact integers, the second stage differs in summation order and in the fp16 hi/lo split of the tensor-core path.  Each
    kernel is held to 1e-4 relative (1e-2 floor) against the ORACLE (tests above and the full-size test below), so against
    each other the bound is the sum, 2e-4 (observed: 1.08e-4 at 900x330 D=70, nearly all of it the shuffle kernel's float
    running sums over the 330-row band); labels agree on > 99.9 % of the pixels"""
    S = pytest.importorskip("stereo_matching_cuda_b200")
    from stereo_matching_cuda_b200 import api

    L, R = synth.make_pair(w, h, max(size_d, 2), channels=3, seed=w + 1)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
    monkeypatch.setenv("SB200_RGB_KERNEL", "4")
    with S.Context(0) as ctx:
        assert ctx.rgb_kernel == 4
        a = ctx.pipeline(L, R, p, want=("disp_left", "disp_right", "best_left", "best_right", "occlusion", "filled"))
    b = _rgb_pipeline_with_kernel(S, api, monkeypatch, 3, L, R, p)
    for kb, kd in (("best_left", "disp_left"), ("best_right", "disp_right")):
        rel = np.abs(a[kb] - b[kb]) / np.maximum(np.abs(b[kb]), 1e-2)
        assert rel.max() < 2e-4, (kb, rel.max())
        assert (a[kd] == b[kd]).mean() > 0.999


@pytest.mark.gpu
def test_gpu_rgb_three_stage_kernel_full_size_1080p_d256(monkeypatch):
    """BASELINE configs[2] at full size: both RGB kernels agree bit for bit on every output map, and the left labels
    recover the synthetic disparity staircase away from the band edges"""
    S = pytest.importorskip("stereo_matching_cuda_b200")
    from stereo_matching_cuda_b200 import api

    w, h, size_d = 1920, 1080, 256
    L, R = synth.make_pair(w, h, size_d, channels=3, seed=3)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
    a = _rgb_pipeline_with_kernel(S, api, monkeypatch, 2, L, R, p)
    b = _rgb_pipeline_with_kernel(S, api, monkeypatch, 3, L, R, p)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32) if a[k].dtype == np.float32 else a[k],
                              b[k].view(np.uint32) if b[k].dtype == np.float32 else b[k]), k
    truth = -synth.delta_rows(h, size_d).astype(np.float32)[:, None]
    inner = np.zeros((h, w), bool)
    inner[:, size_d + 20:w - 20] = True
    ok = (b["disp_left"] == truth) & inner
    assert ok.sum() / inner.sum() > 0.9


@pytest.mark.gpu
def test_gpu_rgb_full_size_1080p_d256_against_the_oracle(oracle):
    """BASELINE configs[2] as specified (RGB guide, 1920x1080, D=256): BOTH views of the default RGB kernel against the
    oracle's exact-mode colour guided filter -- labels identical wherever the oracle's margin is decisive and
    >= 99.9 % overall, best cost within 1e-4 relative (floor 1e-2)."""
    S = pytest.importorskip("stereo_matching_cuda_b200")
    from stereo_matching_cuda_b200 import api

    w, h, size_d = 1920, 1080, 256
    L, R = synth.make_pair(w, h, size_d, channels=3, seed=3)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
    with S.Context(0) as ctx:
        out = ctx.pipeline(L, R, p, want=("gray_left", "gray_right", "disp_left", "disp_right", "best_left", "best_right"))
    gl, gr = oracle.rgb_to_gray(L), oracle.rgb_to_gray(R)
    assert np.array_equal(out["gray_left"], gl) and np.array_equal(out["gray_right"], gr)
    po = oracle.params(box_mode=O.BOX_EXACT, nthreads=oracle.max_threads())
    record = {}
    for view, (rgb, g_guide, g_other, dmin, kd, kb) in {"left": (L, gl, gr, -(size_d - 1), "disp_left", "best_left"),
                                                        "right": (R, gr, gl, 0, "disp_right", "best_right")}.items():
        best, dmap, second = oracle.view_disparity_rgb(rgb, g_guide, g_other, size_d, dmin, po, want_second=True)
        same = out[kd] == dmap
        err = np.abs(out[kb] - best)
        rel = err / np.maximum(np.abs(best), 1e-2)
        record[view] = {"label_agreement": float(same.mean()), "max_abs_err": float(err.max()),
                        "max_rel_err_floor1e-2": float(rel.max()), "median_best": float(np.median(best))}
        assert same.mean() >= 0.999, (view, same.mean())
        assert np.all(same[(second - best) > 2e-4]), view
        assert rel.max() < 1e-4, (view, rel.max(), err.max())
    import json
    import os
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "fullsize_parity_rgb.json"), "w") as f:
        json.dump(record, f)


@pytest.mark.gpu
def test_gpu_rgb_strip_mode_matches_whole_frame():
    """Row strips with 2*radius halo rows and the RGB guide: the strips' labels, concatenated, equal the whole frame's
    (first-stage sums are exact; window areas and statistics use frame rows, so only second-stage float order differs)."""
    S = pytest.importorskip("stereo_matching_cuda_b200")
    torch = pytest.importorskip("torch")
    from stereo_matching_cuda_b200 import api

    w, h, size_d = 300, 200, 24
    L, R = synth.make_pair(w, h, size_d, channels=3, seed=11)
    p = api.default_params(dmin=-(size_d - 1), dmax=0, guide_mode=S.GUIDE_RGB)
    with S.Context(0) as ctx:
        whole = ctx.pipeline(L, R, p, want=("disp_left", "disp_right", "filled", "best_left"))
        halo = ctx.strip_halo_rows(p)
        ctx.set_stream(torch.cuda.current_stream())
        parts = {k: [] for k in ("disp_left", "disp_right", "filled", "best_left")}
        bounds = [0, 64, 128, 200]
        for y0, y1 in zip(bounds[:-1], bounds[1:]):
            top, bot = min(halo, y0), min(halo, h - y1)
            dl = torch.from_numpy(L[y0 - top:y1 + bot].copy()).cuda()
            dr = torch.from_numpy(R[y0 - top:y1 + bot].copy()).cuda()
            outs = {k: torch.empty((y1 - y0, w), dtype=torch.float32, device="cuda") for k in parts}
            ctx.pipeline_strip_dev(dl, dr, 3, w, dict(y0=y0, rows=y1 - y0, halo_top=top, halo_bot=bot, frame_h=h), outs, p)
            torch.cuda.synchronize()
            for k in parts:
                parts[k].append(outs[k].cpu().numpy())
    for k in ("disp_left", "disp_right", "filled"):
        assert (np.concatenate(parts[k], 0) == whole[k]).mean() > 0.9999, k
    assert np.allclose(np.concatenate(parts["best_left"], 0), whole["best_left"], rtol=1e-5, atol=1e-6)
