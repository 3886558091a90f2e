"""Pins oracle/stereo_oracle.c against the reference's own CPU twins: live against
oracle/_ref/libref.so (compiled from the unmodified sources) when it is built, and against
tests/golden/ref_twins_small.npz (outputs of that library, see make_fixtures.py) always."""
import os

import numpy as np
import pytest

import _oracle as O


@pytest.fixture(scope="module")
def fx():
    return dict(np.load(os.path.join(O.GOLDEN_DIR, "ref_twins_small.npz")))


def test_fixture_stage_outputs(oracle, fx):
    size_d, dmin = int(fx["size_d"]), int(fx["dmin"])
    assert np.array_equal(oracle.rgb_to_gray(fx["rgb"]), fx["gray"])
    assert np.array_equal(oracle.x_derivative(fx["left"]), fx["grad_left"])
    assert np.array_equal(oracle.cost_volume(fx["left"], fx["right"], size_d, dmin), fx["cost"])
    fimg = fx["left"].astype(np.float32) * 1.5
    sat = oracle.integral(fimg)
    assert np.array_equal(sat, fx["sat"])
    assert np.array_equal(oracle.box_from_sat(sat), fx["box"])


def test_fixture_pipeline(oracle, fx):
    size_d, dmin = int(fx["size_d"]), int(fx["dmin"])
    r = oracle.pipeline_gray(fx["left"], fx["right"], dmin, size_d, oracle.params(use_fma=0))
    # the CPU twin divides by (var+EPS) in double where the device multiplies by a rounded
    # reciprocal (guidedFilter.cu:608 vs :350): best costs agree to ~1e-4 relative, labels
    # agree except where the WTA margin is below that
    for k in ("bestL", "bestR"):
        assert np.allclose(r[k], fx[k], rtol=2e-4, atol=1e-6)
    for k in ("dL", "dR"):
        assert (r[k] == fx[k]).mean() > 0.995
    # occlusion + fill are exact functions of the labels: check them on the fixture's labels
    occ = oracle.detect_occlusion(fx["dL"], fx["dR"], dmin - 100)
    assert np.array_equal(occ, fx["occ"])
    assert np.array_equal(oracle.fill_occlusion(occ, dmin), fx["filled"])


def test_live_cpu_twins(oracle, reflib):
    rng = np.random.default_rng(7)
    for (w, h) in ((33, 21), (64, 40), (19, 19), (7, 50)):
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        b = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(oracle.rgb_to_gray(rgb), reflib.rgb_to_gray_cpu(rgb))
        assert np.array_equal(oracle.x_derivative(a), reflib.x_derivative_cpu(a))
        for dmin, size_d in ((-4, 6), (0, 5), (2, 3)):
            assert np.array_equal(oracle.cost_volume(a, b, size_d, dmin), reflib.cost_volume_cpu(a, b, size_d, dmin))
        f = rng.random((h, w), dtype=np.float32) * 300
        sat = reflib.integral_cpu(f)
        assert np.array_equal(oracle.integral(f), sat)
        assert np.array_equal(oracle.box_from_sat(sat), reflib.box_filter_cpu(f, sat))
        d = rng.integers(-6, 1, (h, w)).astype(np.float32)
        d[rng.random((h, w)) < 0.3] = -106
        assert np.array_equal(oracle.fill_occlusion(d, -6), reflib.fill_occlusion_cpu(d, -6))


def test_live_detect_occlusion(oracle, reflib):
    rng = np.random.default_rng(8)
    h, w = 30, 50
    dL = rng.integers(-9, 1, (h, w)).astype(np.float32)
    dR = rng.integers(0, 10, (h, w)).astype(np.float32)
    assert np.array_equal(oracle.detect_occlusion(dL, dR, -109), reflib.detect_occlusion_cpu(dL, dR, -109))


def test_live_guided_filter_best_cost(oracle, reflib):
    """guided_filter_onCpu (guidedFilter.cu:540-653): its best-cost output is valid (its dmap
    is not, :622).  It uses /(var+EPS) in double, so compare within 2e-4 relative."""
    rng = np.random.default_rng(9)
    h, w, size_d, dmin = 40, 56, 4, -3
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    other = np.roll(img, 2, axis=1)
    cost = oracle.cost_volume(img, other, size_d, dmin)
    best_ref, _, mean_ref = reflib.guided_filter_cpu(img, cost, dmin, oracle.best_init())
    best, _, mean, _ = oracle.guided_filter(img, cost, dmin, oracle.params(use_fma=0))
    assert np.allclose(best, best_ref, rtol=2e-4, atol=1e-6)
    assert np.array_equal(mean, mean_ref)


def test_threads_do_not_change_results(oracle):
    rng = np.random.default_rng(10)
    h, w = 40, 64
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = np.roll(a, 3, axis=1)
    r1 = oracle.pipeline_gray(a, b, -7, 8, oracle.params(nthreads=1))
    r4 = oracle.pipeline_gray(a, b, -7, 8, oracle.params(nthreads=4))
    for k in ("dL", "dR", "occ", "filled", "bestL", "bestR"):
        assert np.array_equal(r1[k], r4[k])


def test_tie_break_last_slice_wins(oracle):
    """guidedFilter.cu:406 `best >= q`: on a constant pair every in-range slice filters to the
    same q, so the label is the LAST slice (SURVEY.md App. C test 6)."""
    img = np.full((30, 40), 77, np.uint8)
    best, dmap, _, _ = oracle.view_disparity(img, img, 5, -4)
    # left view: d=0 (last slice) is in range everywhere and q==0 there
    assert np.all(dmap == 0)
    q = np.zeros((4, 4), np.float32)
    b = np.full((4, 4), oracle.best_init(), np.float32)
    d = np.zeros((4, 4), np.float32)
    oracle.disp_select(q, b, d, 3)
    oracle.disp_select(q, b, d, 5)
    assert np.all(d == 5)
