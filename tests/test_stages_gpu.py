"""GPU tests of the stand-alone mirrors of dense_lucaskanade's helper functions
(pysteps_b200.stages) against the oracle restatements of the same reference functions."""
import numpy as np
import pytest
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stages():
    import torch
    assert torch.cuda.is_available()
    from pysteps_b200 import stages
    return stages


def _frames():
    from lk_cases import build_case
    return build_case("nan_200x176")[0]


def test_morph_detection_tracking(stages):
    from oracle import lucaskanade as ora
    fr = _frames()
    a = np.ma.masked_invalid(fr[0]); np.ma.set_fill_value(a, a.min())
    b = np.ma.masked_invalid(fr[1]); np.ma.set_fill_value(b, b.min())
    oa, ob = ora.morph_opening(a, a.min(), 3), ora.morph_opening(b, b.min(), 3)
    ga, gb = stages.morph_opening(a, a.min(), 3), stages.morph_opening(b, b.min(), 3)
    keep = ~np.ma.getmaskarray(oa)
    assert np.array_equal(np.ma.getmaskarray(ga), np.ma.getmaskarray(oa))
    assert np.array_equal(ga.data[keep], oa.data[keep])
    # ndarray in -> ndarray out (utils/images.py:83-84)
    plain = np.nan_to_num(fr[0], nan=0.0)
    assert np.array_equal(stages.morph_opening(plain, 0.0, 3), ora.morph_opening(plain, 0.0, 3))
    pts = stages.detection(oa)
    assert_bits_equal(pts.astype(np.float32), ora.detection(oa).astype(np.float32), "corners")
    pts5 = stages.detection(oa, max_num_features=25, min_distance=15, quality_level=0.02)
    assert_bits_equal(pts5.astype(np.float32),
                      ora.detection(oa, max_num_features=25, min_distance=15, quality_level=0.02).astype(np.float32),
                      "corners kwargs")
    xy, uv = stages.track_features(oa, ob, pts.astype(np.float32))
    oxy, ouv = ora.track_features(oa, ob, pts.astype(np.float32))
    assert_bits_equal(np.asarray(xy, np.float32), np.asarray(oxy, np.float32), "xy")
    assert_bits_equal(np.asarray(uv, np.float32), np.asarray(ouv, np.float32), "uv")
    xy2, uv2 = stages.track_features(oa, ob, pts[:40].astype(np.float32), winsize=(21, 21), nr_levels=2)
    o2 = ora.track_features(oa, ob, pts[:40].astype(np.float32), winsize=(21, 21), nr_levels=2)
    assert_bits_equal(np.asarray(uv2, np.float32), np.asarray(o2[1], np.float32), "uv small window")
    e_xy, e_uv = stages.track_features(oa, ob, np.empty((0, 2), np.float32))
    assert e_xy.shape == (0, 2) and e_uv.shape == (0, 2)
    assert stages.detection(np.zeros((64, 64))).shape == (0, 2)


def test_cleansing_and_interpolation(stages):
    # pysteps/tests/test_utils_cleansing.py and test_utils_interpolate.py semantics
    from oracle import lucaskanade as ora
    rng = np.random.default_rng(5)
    xy = np.floor(rng.uniform(0, 300, (700, 2)))
    uv = np.stack([2 + 0.2 * rng.standard_normal(700), -1 + 0.2 * rng.standard_normal(700)], 1)
    uv[::40] += 5.0
    got = stages.detect_outliers(uv, 3, xy, 30)
    want = ora.detect_outliers(uv, 3, xy, 30)
    assert np.array_equal(got, want) and got.sum() > 5
    assert not stages.detect_outliers(uv[:1], 3, xy[:1], 30).any()
    dxy, duv = stages.decluster(xy[~got], uv[~got], 20, 1)
    oxy, ouv = ora.decluster(xy[~want], uv[~want], 20, 1)
    assert np.array_equal(dxy, oxy) and np.array_equal(duv, ouv)
    d3 = stages.decluster(xy, uv, 50, 3)
    o3 = ora.decluster(xy, uv, 50, 3)
    assert np.array_equal(d3[0], o3[0]) and np.array_equal(d3[1], o3[1])
    xg, yg = np.arange(211), np.arange(190)
    for kw in ({}, {"k": 5, "power": 2.0}, {"dist_offset": 1.0, "k": 32}):
        g = stages.idwinterp2d(dxy, duv, xg, yg, **kw)
        o = ora.idwinterp2d(oxy, ouv, xg, yg, **kw)
        assert g.shape == o.shape and np.abs(g - o).max() <= 1e-11, kw
    # one variable, one sample, uniform values (test_utils_interpolate.py)
    g1 = stages.idwinterp2d(dxy, duv[:, 0], xg, yg)
    assert g1.shape == (190, 211) and np.abs(g1 - ora.idwinterp2d(oxy, ouv[:, 0], xg, yg)).max() <= 1e-11
    one = stages.idwinterp2d(dxy[:1], duv[:1], xg, yg)
    assert one.shape == (2, 190, 211) and np.all(one[0] == duv[0, 0]) and np.all(one[1] == duv[0, 1])
    uni = stages.idwinterp2d(dxy, np.full((len(dxy), 2), 3.5), xg, yg)
    assert np.all(uni == 3.5)
    with pytest.raises(ValueError, match="non-finite"):
        stages.idwinterp2d(dxy, np.where(duv > 2.3, np.nan, duv), xg, yg)
    with pytest.raises(ValueError, match="does not match"):
        stages.idwinterp2d(dxy[:-1], duv, xg, yg)
    # a row band of the grid equals the rows of the full fill (tile partitioning of the fill)
    full = stages.idwinterp2d(dxy, duv, xg, yg)
    band = stages.idwinterp2d(dxy, duv, xg, yg[60:130])
    assert np.array_equal(band, full[:, 60:130])


def test_global_outlier_test_and_all_point_weighting(env=None):
    """The k=None branches: cleansing.py:201-214 (every vector against the mean / covariance of all)
    and interpolate.py:82-88 (every vector weighs in at every grid point), stand-alone and inside
    dense_lucaskanade."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a GPU"
    from oracle import lucaskanade as ora
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200 import stages
    from pysteps_b200.motion import get_method
    rng = np.random.default_rng(4)
    for n in (2, 3, 50, 700, 2500):
        uv = np.stack([2 + 0.4 * rng.standard_normal(n), -1 + 0.3 * rng.standard_normal(n)], 1)
        uv[::9] += 3.0
        for thr in (1.0, 2.0, 3.0):
            got = stages.detect_outliers(uv, thr)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = ora.detect_outliers(uv, thr)
            assert np.array_equal(got, want), (n, thr)
    for npts, (ny, nx) in ((3, (20, 30)), (40, (64, 70)), (1500, (90, 100))):
        xy = rng.uniform(0, 100, (npts, 2))
        vals = rng.standard_normal((npts, 2))
        gx, gy = np.arange(nx, dtype=float), np.arange(ny, dtype=float)
        got = stages.idwinterp2d(xy, vals, gx, gy, k=None)
        want = ora.idwinterp2d(xy, vals, gx, gy, k=None)
        assert np.abs(got - want).max() <= 1e-12, npts
    fr = syn.rain_frames(160, 192, 2, 1)
    lk = get_method("lk")
    with ora.knn_mode("ckdtree"):
        xy, uv = lk(fr, dense=False, k_outlier=None)
        oxy, ouv = ora.dense_lucaskanade(fr, dense=False, k_outlier=None)
        assert np.array_equal(xy, oxy) and np.array_equal(uv, ouv)
        V = lk(fr, interp_kwargs={"k": None})
        Vo = ora.dense_lucaskanade(fr, interp_kwargs={"k": None})
        assert np.abs(V - Vo).max() <= 1e-12
