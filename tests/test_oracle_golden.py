"""Pins oracle/stereo_oracle.c against the reference's 12 golden PNGs (SURVEY.md section 4.3,
written by main.cu:162-181 from a Tsukuba run, D_MIN=-15, D_MAX=0).  CPU only."""
import numpy as np
import pytest

import _oracle as O

DMIN, SIZE_D = -15, 16


@pytest.fixture(scope="module")
def tsukuba(oracle):
    L, R = O.tsukuba_rgb()
    gl, gr = oracle.rgb_to_gray(L), oracle.rgb_to_gray(R)
    p = oracle.params(box_mode=O.BOX_FAITHFUL, use_fma=0, nthreads=oracle.max_threads())
    res = oracle.pipeline_gray(gl, gr, DMIN, SIZE_D, p, want_second=True)
    res["gl"], res["gr"] = gl, gr
    return res


def test_gray_exact(tsukuba):
    assert np.array_equal(tsukuba["gl"], O.load_png("image_left.png"))
    assert np.array_equal(tsukuba["gr"], O.load_png("image_right.png"))


def test_mean_image_exact(tsukuba):
    # (uchar)(int)box_r9(I): pins the float32 SAT box mean incl. border clipping
    assert np.array_equal(tsukuba["meanL"], O.load_png("image_mean_left.png"))
    assert np.array_equal(tsukuba["meanR"], O.load_png("image_mean_right.png"))


def test_cost_slice0(oracle, tsukuba):
    cl = oracle.cost_volume(tsukuba["gl"], tsukuba["gr"], 1, DMIN)[0]
    cr = oracle.cost_volume(tsukuba["gr"], tsukuba["gl"], 1, 0)[0]
    assert np.array_equal(oracle.write_mat(cl), O.load_png("cost_lminus15.png"))
    assert np.array_equal(oracle.write_mat(cr), O.load_png("cost_rminus15.png"))


def test_best_cost(oracle, tsukuba):
    # min/max-normalised to 8 bit by write_mat, so every pixel depends on the last bit of the
    # global min and max: an exact match pins the whole float32 filter chain
    for key, name in (("bestL", "best_costl.png"), ("bestR", "best_costr.png")):
        assert np.array_equal(oracle.write_mat(tsukuba[key]), O.load_png(name))


def test_disparity_labels_exact(tsukuba):
    gl = O.load_png("disparity_mapl.png").astype(int)
    gr = O.load_png("disparity_mapr.png").astype(int)
    assert np.array_equal(tsukuba["dL"].astype(int), gl // 17 - 15)
    assert np.array_equal(tsukuba["dR"].astype(int), gr // 17)
    assert np.all(gl % 17 == 0) and np.all(gr % 17 == 0)


def test_occlusion_and_fill_exact(tsukuba):
    occ_png = O.load_png("occlu_mapl.png").astype(int)
    occ = tsukuba["occ"]
    assert np.array_equal(occ == DMIN - 100, occ_png == 0)
    assert abs((occ_png == 0).mean() - 0.09589) < 1e-4
    # non-occluded pixels: write_mat maps d in [-115, 0] to (d+115)*255/115
    keep = occ_png != 0
    assert np.array_equal(((occ[keep] + 115.0) * 255.0 / 115.0).astype(int), occ_png[keep])
    filled_png = O.load_png("occlu_mapl_filled.png").astype(int)
    assert np.array_equal(tsukuba["filled"].astype(int), filled_png // 17 - 15)


def test_exact_mode_close_to_faithful(oracle, tsukuba):
    """H1 in SURVEY.md: the float32 SAT is noisier than a true box sum; labels still agree
    everywhere the reference's own WTA margin exceeds that noise."""
    p = oracle.params(box_mode=O.BOX_EXACT, use_fma=0, nthreads=oracle.max_threads())
    ex = oracle.pipeline_gray(tsukuba["gl"], tsukuba["gr"], DMIN, SIZE_D, p)
    for lab, best, second in (("dL", "bestL", "secondL"), ("dR", "bestR", "secondR")):
        same = ex[lab] == tsukuba[lab]
        assert same.mean() > 0.9998
        margin = tsukuba[second] - tsukuba[best]
        assert np.all(same[margin > 5e-3])
        rel = np.abs(ex[best] - tsukuba[best]) / np.maximum(np.abs(tsukuba[best]), 1e-3)
        assert np.median(rel) < 1e-4


def test_fma_switch_is_label_neutral(oracle, tsukuba):
    p = oracle.params(box_mode=O.BOX_FAITHFUL, use_fma=1, nthreads=oracle.max_threads())
    nf = oracle.pipeline_gray(tsukuba["gl"], tsukuba["gr"], DMIN, SIZE_D, p)
    assert (nf["dL"] == tsukuba["dL"]).mean() > 0.9999
    assert np.array_equal(oracle.cost_volume(tsukuba["gl"], tsukuba["gr"], 4, -3, oracle.params(use_fma=0)),
                          oracle.cost_volume(tsukuba["gl"], tsukuba["gr"], 4, -3, oracle.params(use_fma=1)))


def test_oracle_weighted_median_properties(oracle):
    """the beyond-reference weighted median (stereo_b200.h): unmarked pixels pass through, a marked pixel in a constant
    neighbourhood keeps that constant, and with a huge colour sigma and radius 1 it is the plain median of the 3x3 window"""
    rng = np.random.default_rng(7)
    h, w, dmin, size_d = 20, 30, -7, 8
    gray = rng.integers(0, 255, (h, w)).astype(np.uint8)
    filled = rng.integers(dmin, 1, (h, w)).astype(np.float32)
    occ = filled.copy()
    marked = rng.random((h, w)) < 0.3
    occ[marked] = dmin - 100
    out = oracle.weighted_median(gray, occ, filled, dmin, size_d, radius=1, sigma_space=1e6, sigma_color=1e6)
    assert np.array_equal(out[~marked], filled[~marked])
    for y, x in zip(*np.nonzero(marked)):
        win = np.sort(filled[max(0, y - 1):y + 2, max(0, x - 1):x + 2].ravel())
        # lower weighted median with equal weights: first value whose count reaches half
        k = next(i for i in range(len(win)) if 2 * (i + 1) >= len(win))
        assert out[y, x] == win[k]
    const = np.full((h, w), -3, np.float32)
    out2 = oracle.weighted_median(gray, occ, const, dmin, size_d)
    assert np.array_equal(out2, const)


def test_oracle_subpixel_refine_against_numpy(oracle):
    """so_subpixel_refine (beyond the reference; definition in stereo_b200.h) against an independent numpy
    restatement in float32 with the same operation order, on a random volume: bit for bit; plus the closed form on an exact
    parabola and the rules at the edges of the label range and on marked pixels"""
    rng = np.random.default_rng(5)
    size_d, h, w, dmin = 9, 13, 17, -8
    vol = rng.random((size_d, h, w), dtype=np.float32)
    k = vol.argmin(axis=0)
    disp = (k + dmin).astype(np.float32)
    got = oracle.subpixel_refine(vol, disp, dmin)
    f = np.float32
    qm = np.take_along_axis(vol, np.clip(k - 1, 0, size_d - 1)[None], 0)[0]
    q0 = np.take_along_axis(vol, k[None], 0)[0]
    qp = np.take_along_axis(vol, np.clip(k + 1, 0, size_d - 1)[None], 0)[0]
    den = (qm - q0).astype(f) + (qp - q0).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.clip(((f(0.5) * (qm - qp).astype(f)).astype(f) / den).astype(f), f(-0.5), f(0.5))
    want = np.where((k > 0) & (k < size_d - 1) & (den > 0), (disp + t).astype(f), disp)
    assert np.array_equal(got.view(np.uint32), want.astype(f).view(np.uint32))
    # an exact parabola with its vertex at -3.25: q(d) = (d + 3.25)^2
    d = np.arange(dmin, dmin + size_d, dtype=np.float32)
    par = np.broadcast_to(((d + 3.25) ** 2)[:, None, None], (size_d, 2, 3)).astype(np.float32).copy()
    r = oracle.subpixel_refine(par, np.full((2, 3), -3.0, np.float32), dmin)
    assert np.allclose(r, -3.25, atol=1e-6)
    # edge labels stay; marked pixels take the filled label
    edge = oracle.subpixel_refine(par, np.full((2, 3), float(dmin), np.float32), dmin)
    assert np.array_equal(edge, np.full((2, 3), float(dmin), np.float32))
    occ = np.full((2, 3), float(dmin - 100), np.float32)
    fil = np.full((2, 3), -5.0, np.float32)
    assert np.array_equal(oracle.subpixel_refine(par, np.full((2, 3), -3.0, np.float32), dmin, occ, fil), fil)
