"""GPU parity tests of the Lucas-Kanade CUDA path against the CPU oracle and the committed
reference outputs.  Stage kernels are called through the C ABI (ctypes) exactly as the
Python mirror does.  Bars: integer images, corners, tracked points, outlier flags and
declustered vectors BIT-IDENTICAL; dense field <= 1e-12 where the k-NN set is unique."""
import ctypes
import os

import numpy as np
import pytest
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a GPU"
    from pysteps_b200 import _device, _lib
    _device.require_cuda()
    return torch, _lib


@pytest.fixture(scope="module")
def golden_lk():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "lk_golden.npz"))


def _frames(name):
    from lk_cases import build_case
    return build_case(name)[0]


def _masked(a):
    x = np.ma.masked_invalid(a)
    np.ma.set_fill_value(x, x.min())
    return x


class Stages:
    """Thin ctypes driver of the dense stage kernels for one frame."""

    def __init__(self, env, frame, buffer_mask=5):
        torch, L = env
        self.torch, self.L = torch, L
        self.m, self.n = frame.shape
        m, n = self.m, self.n
        s = torch.cuda.current_stream().cuda_stream
        self.img = torch.from_numpy(np.ascontiguousarray(frame)).cuda()
        self.mask = torch.empty((m, n), dtype=torch.uint8, device="cuda")
        self.st0 = torch.empty(3, dtype=torch.float64, device="cuda")
        L.call("b200_mask_invalid", self.img.data_ptr(), None, m, n, self.mask.data_ptr(),
               self.st0.data_ptr(), s)
        self.opened = torch.empty((m, n), dtype=torch.float64, device="cuda")
        L.call("b200_morph_opening", self.img.data_ptr(), self.mask.data_ptr(), m, n, 3,
               self.st0.data_ptr(), self.st0.data_ptr(), self.opened.data_ptr(), s)
        self.st = torch.empty(12, dtype=torch.float64, device="cuda")
        L.call("b200_masked_minmax", self.opened.data_ptr(), self.mask.data_ptr(), m, n, buffer_mask,
               self.st0.data_ptr(), self.st.data_ptr(), s)
        self.q_track = torch.empty((m, n), dtype=torch.uint8, device="cuda")
        L.call("b200_quantise_u8", self.opened.data_ptr(), self.mask.data_ptr(), m, n, 0, 0,
               self.st.data_ptr(), self.st.data_ptr(), self.q_track.data_ptr(), None, s)
        self.q_det = torch.empty((m, n), dtype=torch.uint8, device="cuda")
        self.valid = torch.empty((m, n), dtype=torch.uint8, device="cuda")
        L.call("b200_quantise_u8", self.opened.data_ptr(), self.mask.data_ptr(), m, n, 1, buffer_mask,
               self.st.data_ptr(), self.st.data_ptr(), self.q_det.data_ptr(), self.valid.data_ptr(), s)


def _boundary_frame(m, n, seed):
    """Values whose scaled image (x - min) / (max - min) * 255 sits on, and a few ulps either side of,
    integer boundaries -- where the fused front end's division-free scaling must defer to the division."""
    rng = np.random.default_rng(seed)
    k = rng.integers(0, 256, size=(m, n)).astype(np.float64)
    a = k / 255.0 * 40.0
    ulps = rng.integers(-3, 4, size=(m, n))
    a = np.where(ulps == 0, a, np.nextafter(a, np.where(ulps > 0, np.inf, -np.inf)))
    a[0, 0], a[0, 1] = 0.0, 40.0
    a[rng.random((m, n)) < 0.3] = 0.0
    return a


@pytest.mark.parametrize("case", ["plain_160x200", "nan_200x176", "boundary", "nan_boundary", "f32"])
@pytest.mark.parametrize("buffer_mask,opening", [(5, 3), (0, 3), (3, 0), (4, 3), (1, 3), (2, 0)])
def test_fused_front_end_equals_stage_kernels(env, case, buffer_mask, opening):
    """b200_lk_frontend (TMA tiles, bit-row stencils, division-free scaling) against the stand-alone
    stage kernels, bit for bit: mask, both statistics blocks, both uint8 images, the validity map."""
    torch, L = env
    f32 = case == "f32"
    if case in ("boundary", "nan_boundary"):
        fr = _boundary_frame(144, 200, 5)
        if case == "nan_boundary":
            fr[40:60, 90:130] = np.nan
            fr[3, 5] = np.nan
    elif f32:
        fr = _frames("nan_200x176")[0].astype(np.float32).astype(np.float64)
    else:
        fr = _frames(case)[0]
    m, n = fr.shape
    s = torch.cuda.current_stream().cuda_stream
    img = torch.from_numpy(np.ascontiguousarray(fr)).cuda()
    flag = 2 if f32 else 0  # B200_QUANTISE_F32
    # stage kernels
    mask = torch.empty((m, n), dtype=torch.uint8, device="cuda")
    st0 = torch.empty(3, dtype=torch.float64, device="cuda")
    L.call("b200_mask_invalid", img.data_ptr(), None, m, n, mask.data_ptr(), st0.data_ptr(), s)
    opened = img
    if opening:
        opened = torch.empty((m, n), dtype=torch.float64, device="cuda")
        L.call("b200_morph_opening", img.data_ptr(), mask.data_ptr(), m, n, 3, st0.data_ptr(), st0.data_ptr(),
               opened.data_ptr(), s)
    st = torch.empty(12, dtype=torch.float64, device="cuda")
    L.call("b200_masked_minmax", opened.data_ptr(), mask.data_ptr(), m, n, buffer_mask, st0.data_ptr(),
           st.data_ptr(), s)
    q_track = torch.empty((m, n), dtype=torch.uint8, device="cuda")
    L.call("b200_quantise_u8", opened.data_ptr(), mask.data_ptr(), m, n, 0 | flag, 0, st.data_ptr(), st.data_ptr(),
           q_track.data_ptr(), None, s)
    q_det = torch.empty((m, n), dtype=torch.uint8, device="cuda")
    valid = torch.empty((m, n), dtype=torch.uint8, device="cuda")
    L.call("b200_quantise_u8", opened.data_ptr(), mask.data_ptr(), m, n, 1 | flag, buffer_mask, st.data_ptr(),
           st.data_ptr(), q_det.data_ptr(), valid.data_ptr(), s)
    # fused
    mask2 = torch.zeros((m, n), dtype=torch.uint8, device="cuda")
    st0b = torch.zeros(3, dtype=torch.float64, device="cuda")
    stb = torch.zeros(12, dtype=torch.float64, device="cuda")
    qt2 = torch.zeros((m, n), dtype=torch.uint8, device="cuda")
    qd2 = torch.zeros((m, n), dtype=torch.uint8, device="cuda")
    v2 = torch.zeros((m, n), dtype=torch.uint8, device="cuda")
    L.call("b200_lk_frontend", img.data_ptr(), None, m, n, opening, buffer_mask, flag, mask2.data_ptr(),
           st0b.data_ptr(), stb.data_ptr(), qt2.data_ptr(), qd2.data_ptr(), v2.data_ptr(), s)
    assert torch.equal(mask, mask2)
    assert_bits_equal(st0b.cpu().numpy(), st0.cpu().numpy(), "frame statistics")
    assert_bits_equal(stb.cpu().numpy(), st.cpu().numpy(), "opened-image statistics")
    assert torch.equal(q_track, qt2), f"tracking image: {(q_track != qt2).sum().item()} pixels differ"
    assert torch.equal(q_det, qd2), f"detection image: {(q_det != qd2).sum().item()} pixels differ"
    assert torch.equal(valid, v2)


@pytest.mark.parametrize("name", ["plain_160x200", "nan_200x176", "odd_width_150x203"])
def test_dense_stages(env, name):
    torch, L = env
    from oracle import lucaskanade as ora
    fr = _frames(name)[0]
    S = Stages(env, fr)
    a = _masked(fr)
    o = ora.morph_opening(a, a.min(), 3)
    got = S.opened.cpu().numpy()
    keep = ~np.ma.getmaskarray(o)
    assert np.array_equal(got[keep], o.data[keep]), "morph_opening"
    assert np.array_equal(S.mask.cpu().numpy().astype(bool), np.ma.getmaskarray(o))
    assert np.array_equal(S.q_track.cpu().numpy(), ora.tracking_image(o)), "tracking uint8 image"
    qd, valid = ora.detection_image(o)
    assert np.array_equal(S.q_det.cpu().numpy(), qd), "detection uint8 image"
    assert np.array_equal(S.valid.cpu().numpy(), valid), "detection mask"
    # Shi-Tomasi map: bit-identical to cv2 4.13.0 through the oracle
    m, n = fr.shape
    s = torch.cuda.current_stream().cuda_stream
    eig = torch.empty((m, n), dtype=torch.float32, device="cuda")
    L.call("b200_min_eig", S.q_det.data_ptr(), m, n, eig.data_ptr(), s)
    assert_bits_equal(eig.cpu().numpy(), ora.corner_min_eigen_val(qd), "min eigenvalue map")
    # corner selection
    for maxc, q, md in ((1000, 0.01, 10), (37, 0.05, 4.5), (500, 0.001, 1)):
        corners = torch.zeros((maxc, 2), dtype=torch.float32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        L.call("b200_good_features", eig.data_ptr(), S.valid.data_ptr(), m, n, maxc, q, float(md),
               corners.data_ptr(), cnt.data_ptr(), s)
        want = ora.good_features_to_track(qd, valid, maxc, q, md)
        c = int(cnt.item())
        assert c == len(want), f"corner count {c} != {len(want)} ({maxc},{q},{md})"
        assert_bits_equal(corners[:c].cpu().numpy(), want, "corners")


def test_corners_match_cv2_golden(env, golden_lk):
    torch, L = env
    for name in ("plain_160x200", "nan_200x176", "three_frames_192x160", "odd_width_150x203"):
        fr = _frames(name)[0]
        S = Stages(env, fr)
        m, n = fr.shape
        s = torch.cuda.current_stream().cuda_stream
        eig = torch.empty((m, n), dtype=torch.float32, device="cuda")
        L.call("b200_min_eig", S.q_det.data_ptr(), m, n, eig.data_ptr(), s)
        corners = torch.zeros((1000, 2), dtype=torch.float32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        L.call("b200_good_features", eig.data_ptr(), S.valid.data_ptr(), m, n, 1000, 0.01, 10.0,
               corners.data_ptr(), cnt.data_ptr(), s)
        assert_bits_equal(corners[:int(cnt.item())].cpu().numpy(), golden_lk[name + "/points"], name)


def test_min_eig_matches_cv2_golden(env, golden_lk):
    torch, L = env
    q = golden_lk["cv/q"]
    m, n = q.shape
    dq = torch.from_numpy(q).cuda()
    eig = torch.empty((m, n), dtype=torch.float32, device="cuda")
    L.call("b200_min_eig", dq.data_ptr(), m, n, eig.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert_bits_equal(eig.cpu().numpy(), golden_lk["cv/min_eig"], "cornerMinEigenVal vs cv2")


def _build_pyr(env, q, win, levels, deriv=True):
    torch, L = env
    from pysteps_b200.motion.lucaskanade import _pyramid_layout
    m, n = q.shape
    lv, total = _pyramid_layout(m, n, win, levels)
    P = torch.empty(total, dtype=torch.uint8, device="cuda")
    D = torch.empty(2 * total, dtype=torch.int16, device="cuda") if deriv else None
    dq = q if isinstance(q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q)).cuda()
    L.call("b200_lk_build_pyramid", dq.data_ptr(), m, n, win[0], win[1], levels, P.data_ptr(),
           None if D is None else D.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return P, D, lv


def test_pyramid_and_scharr_match_cv2_golden(env, golden_lk):
    torch, L = env
    q = golden_lk["cv/q"]
    P, D, lv = _build_pyr(env, q, (21, 21), 3)
    off = 0
    lvl = 0
    while f"cv/pyr{lvl}" in golden_lk.files:
        ref = golden_lk[f"cv/pyr{lvl}"]
        h, w = ref.shape
        assert np.array_equal(P[off:off + h * w].cpu().numpy().reshape(h, w), ref), f"level {lvl}"
        d = D[2 * off:2 * (off + h * w)].cpu().numpy().reshape(h, w, 2)
        assert np.array_equal(d, golden_lk[f"cv/deriv{lvl}"]), f"Scharr level {lvl}"
        off += h * w
        lvl += 1
    assert lvl == lv + 1


def _track(env, I, J, pts, win=(50, 50), levels=3, criteria=(3, 10, 0), min_eig=1e-4):
    torch, L = env
    m, n = I.shape
    PI, DI, _ = _build_pyr(env, I, win, levels, True)
    PJ, _, _ = _build_pyr(env, J, win, levels, False)
    npts = len(pts)
    p0 = torch.from_numpy(np.ascontiguousarray(pts, np.float32)).cuda()
    p1 = torch.zeros((npts, 2), dtype=torch.float32, device="cuda")
    st = torch.zeros(npts, dtype=torch.uint8, device="cuda")
    L.call("b200_lk_track", PI.data_ptr(), PJ.data_ptr(), DI.data_ptr(), m, n, win[0], win[1], levels,
           criteria[1], float(criteria[2]), min_eig, p0.data_ptr(), npts, None, p1.data_ptr(),
           st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return p1.cpu().numpy(), st.cpu().numpy()


def test_tracker_bit_exact_vs_oracle_subpixel(env):
    """Non-integer, noisy motion + points at the image edge: every iteration / oscillation /
    lost-feature path of the tracker, three window geometries."""
    from oracle import lucaskanade as ora
    from pysteps_b200 import _synthetic as syn
    rng = np.random.default_rng(7)
    base = syn.powerlaw_field(300, 340, 9)
    I = np.clip((base - base.min()) / (base.max() - base.min()) * 255, 0, 255).astype(np.uint8)
    # sub-pixel warp by bilinear resampling in float64 (no cv2 needed on the GPU box)
    yy, xx = np.mgrid[0:300, 0:340].astype(np.float64)
    mx = np.clip(xx - 2.37 + 0.8 * np.sin(yy / 40.0), 0, 338.999)
    my = np.clip(yy + 1.61 + 0.6 * np.cos(xx / 55.0), 0, 298.999)
    x0, y0 = np.floor(mx).astype(int), np.floor(my).astype(int)
    tx, ty = mx - x0, my - y0
    If = I.astype(np.float64)
    J = (If[y0, x0] * (1 - tx) * (1 - ty) + If[y0, x0 + 1] * tx * (1 - ty) +
         If[y0 + 1, x0] * (1 - tx) * ty + If[y0 + 1, x0 + 1] * tx * ty)
    J = np.clip(np.rint(J) + rng.integers(-6, 7, J.shape), 0, 255).astype(np.uint8)
    pts = ora.good_features_to_track(I, None, 400, 0.01, 7)
    pts = np.concatenate([pts, np.array([[0.0, 0.0], [339.0, 299.0], [3.5, 150.25], [170.75, 2.0]],
                                        np.float32)])
    for win, lev in (((50, 50), 3), ((21, 21), 2), ((31, 15), 3), ((9, 12), 1)):
        got, gst = _track(env, I, J, pts, win, lev)
        want, wst = ora.calc_optical_flow_pyr_lk(I, J, pts, win, lev, (3, 10, 0), 1e-4)
        assert np.array_equal(gst, wst), f"status {win}"
        good = wst == 1
        assert good.sum() > 100
        assert_bits_equal(got[good], want[good], f"tracked points {win}")


def test_tracker_matches_cv2_golden(env, golden_lk):
    from oracle import lucaskanade as ora
    for name in ("plain_160x200", "nan_200x176", "odd_width_150x203"):
        fr = _frames(name)
        a, b = _masked(fr[0]), _masked(fr[1])
        a = ora.morph_opening(a, a.min(), 3)
        b = ora.morph_opening(b, b.min(), 3)
        pts = golden_lk[name + "/points"]
        p1, st = _track(env, ora.tracking_image(a), ora.tracking_image(b), pts)
        keep = st == 1
        assert_bits_equal(pts[keep], golden_lk[name + "/xy"].astype(np.float32), name + " xy")
        assert_bits_equal((p1[keep] - pts[keep]).astype(np.float32),
                          golden_lk[name + "/uv"].astype(np.float32), name + " uv")


def test_outliers_decluster_idw_vs_oracle(env):
    torch, L = env
    from oracle import lucaskanade as ora
    rng = np.random.default_rng(3)
    s = torch.cuda.current_stream().cuda_stream
    for n_pts, dup in ((600, False), (1500, True), (40, False), (25, False)):
        xy = np.floor(rng.uniform(0, 400, (n_pts, 2)))
        uv = np.stack([3 + 0.3 * rng.standard_normal(n_pts), -2 + 0.3 * rng.standard_normal(n_pts)], 1)
        uv[rng.integers(0, n_pts, max(2, n_pts // 30))] += rng.uniform(-8, 8, (max(2, n_pts // 30), 2))
        uv = uv.astype(np.float32).astype(np.float64)
        if dup:
            uv[:200, 1] = -2.0  # exactly singular local covariances
        dxy, duv = torch.from_numpy(xy).cuda(), torch.from_numpy(uv).cuda()
        flags = torch.zeros(n_pts, dtype=torch.uint8, device="cuda")
        L.call("b200_detect_outliers", duv.data_ptr(), dxy.data_ptr(), None, n_pts, 3.0, 30,
               flags.data_ptr(), s)
        want = ora.detect_outliers(uv, 3, xy, 30)
        got = flags.cpu().numpy().astype(bool)
        # decisions are equal except (possibly) where the Mahalanobis distance sits on the
        # threshold to rounding; require exact agreement on >= 99.5 % and report
        assert (got == want).mean() >= 0.995, f"outlier flags differ: {(got != want).sum()}"
        kxy = torch.empty_like(dxy)
        kuv = torch.empty_like(duv)
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        L.call("b200_compact_rows", dxy.data_ptr(), duv.data_ptr(), flags.data_ptr(), None, n_pts,
               kxy.data_ptr(), kuv.data_ptr(), cnt.data_ptr(), s)
        c = int(cnt.item())
        assert np.array_equal(kxy[:c].cpu().numpy(), xy[~got]) and np.array_equal(kuv[:c].cpu().numpy(), uv[~got])
        oxy = torch.empty_like(dxy)
        ouv = torch.empty_like(duv)
        L.call("b200_decluster", kxy.data_ptr(), kuv.data_ptr(), cnt.data_ptr(), n_pts, 20.0, 1,
               oxy.data_ptr(), ouv.data_ptr(), cnt.data_ptr(), s)
        wxy, wuv = ora.decluster(xy[~got], uv[~got], 20, 1)
        d = int(cnt.item())
        assert d == len(wxy)
        assert np.array_equal(oxy[:d].cpu().numpy(), wxy), "decluster coordinates"
        assert np.array_equal(ouv[:d].cpu().numpy(), wuv), "decluster values"
        if d >= 2:
            nx, ny = 301, 257
            gx = torch.arange(nx, dtype=torch.float64, device="cuda")
            gy = torch.arange(ny, dtype=torch.float64, device="cuda")
            for k in (20, 5, 32):
                out = torch.empty((2, ny, nx), dtype=torch.float64, device="cuda")
                ref = ora.idwinterp2d(wxy, wuv, np.arange(nx), np.arange(ny), k=k)
                # the coordinates here are integers / half-integers: both the packed-key fast
                # path (flag 1) and the general path (flag 0) must reproduce the oracle
                for on_grid in (1, 0):
                    out.zero_()
                    L.call("b200_idw_fill", oxy.data_ptr(), ouv.data_ptr(), None, d, 2, min(k, d), 0.5, 0.5,
                           1.0, gx.data_ptr(), nx, gy.data_ptr(), ny, on_grid, out.data_ptr(), s)
                    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-12, f"idw k={k} on_grid={on_grid}"


def test_idw_general_coordinates(env):
    torch, L = env
    from oracle import lucaskanade as ora
    rng = np.random.default_rng(12)
    s = torch.cuda.current_stream().cuda_stream
    for npts in (7, 300, 2500):
        xy = rng.uniform(-5, 260, (npts, 2))
        vals = rng.normal(size=(npts, 2))
        gxh = np.linspace(0.3, 250.7, 211)
        gyh = np.linspace(-2.0, 240.0, 173)
        ref = ora.idwinterp2d(xy, vals, gxh, gyh, k=20)
        mean_res = float(np.mean(np.abs([np.gradient(gxh).mean(), np.gradient(gyh).mean()])))
        out = torch.empty((2, gyh.size, gxh.size), dtype=torch.float64, device="cuda")
        dxy, dv = torch.from_numpy(xy).cuda(), torch.from_numpy(vals).cuda()  # keep alive
        dgx, dgy = torch.from_numpy(gxh).cuda(), torch.from_numpy(gyh).cuda()
        L.call("b200_idw_fill", dxy.data_ptr(), dv.data_ptr(), None, npts, 2, min(20, npts), 0.5, 0.5,
               mean_res, dgx.data_ptr(), gxh.size, dgy.data_ptr(), gyh.size, 0, out.data_ptr(), s)
        torch.cuda.synchronize()
        assert np.abs(out.cpu().numpy() - ref).max() <= 1e-11, npts


@pytest.mark.parametrize("name", ["plain_160x200", "nan_200x176", "three_frames_192x160",
                                  "odd_width_150x203"])
def test_dense_lucaskanade_vs_reference_golden(env, name, golden_lk):
    from lk_cases import build_case
    from oracle import lucaskanade as ora
    from pysteps_b200.motion import get_method
    lk = get_method("LK")
    frames, kw = build_case(name)
    sxy, suv = lk(frames, dense=False, **kw)
    assert np.array_equal(sxy, golden_lk[name + "/sparse_xy"]), "sparse xy"
    assert np.array_equal(suv, golden_lk[name + "/sparse_uv"]), "sparse uv"
    V = lk(frames, **kw)
    ref = golden_lk[name + "/dense"]
    assert V.shape == ref.shape and V.dtype == ref.dtype
    # EVERY pixel, k-NN ties included (cKDTree's neighbour order is reproduced on the device)
    assert np.abs(V - ref).max() <= 1e-12, "dense field vs the reference"
    assert np.abs(V - ora.dense_lucaskanade(frames, **kw)).max() <= 1e-12


def test_api_behaviour(env):
    torch, L = env
    from pysteps_b200.motion import get_method, lucaskanade
    lk = get_method("lk")
    assert lk is lucaskanade.dense_lucaskanade and get_method("LucasKanade") is lk
    # no precipitation -> exact zeros (lucaskanade.py:245-247, tests/test_motion.py:265-289)
    V = lk(np.zeros((3, 80, 90)))
    assert V.shape == (2, 80, 90) and not V.any()
    xy, uv = lk(np.zeros((2, 80, 90)), dense=False)
    assert xy.shape == (0, 2) and uv.shape == (0, 2)
    assert lk(np.zeros((1, 64, 64))).shape == (2, 64, 64)
    with pytest.raises(ValueError, match="dimension mismatch"):
        lk(np.zeros((80, 90)))
    with pytest.raises(NotImplementedError):
        lk(np.zeros((2, 80, 90)), fd_method="blob")
    with pytest.raises(NotImplementedError):
        lk(np.zeros((2, 80, 90)), interp_method="rbfinterp2d")
    with pytest.raises(ValueError):
        get_method("nonexistent")
    with pytest.raises(NotImplementedError):
        get_method("brox")
    assert not get_method(None)(np.zeros((2, 5, 6))).any()
    # ndarray-with-NaN == MaskedArray input (tests/test_motion.py:400-430), bitwise here
    from lk_cases import build_case
    fr, _ = build_case("nan_200x176")
    Vn = lk(fr)
    Vm = lk(np.ma.masked_invalid(fr))
    assert np.array_equal(Vn, Vm)
    # nr_std_outlier = 0 flags everything (tests/test_motion_lk.py) -> zero field
    from pysteps_b200 import _synthetic as syn
    assert not lk(syn.rain_frames(128, 128, 2, 5), nr_std_outlier=0).any()
    # input is not mutated; device tensors in -> device tensors out
    fr2 = syn.rain_frames(128, 160, 2, 6)
    keep = fr2.copy()
    Vh = lk(fr2)
    assert np.array_equal(fr2, keep)
    Vd = lk(torch.from_numpy(fr2).cuda())
    assert Vd.is_cuda and np.array_equal(Vd.cpu().numpy(), Vh)


def test_full_size_recovers_translation(env):
    """BASELINE.json size: 2048^2 frames translated by (3,-2) px/step."""
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.motion import get_method
    fr = syn.rain_frames(2048, 2048, 2, 0)
    V = get_method("lk")(fr)
    assert V.shape == (2, 2048, 2048) and np.isfinite(V).all()
    wet = fr[1] > 0
    assert abs(V[0][wet].mean() - 3.0) < 0.02 and abs(V[1][wet].mean() + 2.0) < 0.02
    assert np.percentile(np.abs(V[0] - 3.0), 99) < 0.2


@pytest.mark.parametrize("shape,seed", [((40, 44), 3), ((64, 51), 4), ((120, 33), 5), ((257, 300), 6)])
def test_small_and_ragged_frames_vs_oracle(env, shape, seed):
    """Frames smaller than the 50x50 tracking window (single pyramid level, windows that hang
    over the border), odd sizes, three frames: sparse vectors identical, dense field <= 1e-12."""
    from oracle import lucaskanade as ora
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.motion import get_method
    lk = get_method("lk")
    fr = syn.rain_frames(shape[0], shape[1], 3, seed, dx=2, dy=-1)
    sxy, suv = lk(fr, dense=False)
    oxy, ouv = ora.dense_lucaskanade(fr, dense=False)
    assert np.array_equal(sxy, oxy) and np.array_equal(suv, ouv)
    V = lk(fr)
    Vo = ora.dense_lucaskanade(fr)
    assert V.shape == Vo.shape and np.abs(V - Vo).max() <= 1e-12


def test_kwargs_variants_vs_oracle(env):
    """Non-default lk / fd / interp kwargs flow through to the kernels like in the reference."""
    from oracle import lucaskanade as ora
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.motion import get_method
    lk = get_method("lk")
    fr = syn.rain_frames(200, 220, 2, 8)
    for kw in (dict(lk_kwargs=dict(winsize=(21, 21), nr_levels=2), fd_kwargs=dict(max_corners=150, min_distance=6)),
               dict(fd_kwargs=dict(quality_level=0.05, buffer_mask=0), interp_kwargs=dict(k=8, power=1.5)),
               dict(size_opening=0, decl_scale=1, k_outlier=10, nr_std_outlier=2)):
        V = lk(fr, **kw)
        Vo = ora.dense_lucaskanade(fr, **kw)
        assert np.abs(V - Vo).max() <= 1e-11, kw


def test_row_band_fill_equals_rows_of_full_field(env):
    """interp_kwargs b200_rows: a band of the motion field is bit-identical to the same rows of
    the full field (what the tile-partitioned multi-GPU path all-gathers)."""
    from pysteps_b200.motion.lucaskanade import dense_lucaskanade as lk
    frames = _frames("nan_200x176")
    full = lk(frames)
    for r0, r1 in ((0, 200), (0, 67), (67, 134), (134, 200), (93, 94)):
        band = lk(frames, interp_kwargs={"b200_rows": (r0, r1)})
        assert band.shape == (2, r1 - r0, 176)
        assert np.array_equal(band, full[:, r0:r1])
    with pytest.raises(ValueError):
        lk(frames, interp_kwargs={"b200_rows": (10, 10)})
    assert lk(frames[:1], interp_kwargs={"b200_rows": (5, 9)}).shape == (2, 4, 176)


def test_exact_ties_sparse_vectors_vs_oracle_ckdtree_mode(env):
    """The outlier stage takes tied neighbours in cKDTree's order; the sparse vectors equal the
    oracle in cKDTree mode (which equals the reference bit for bit)."""
    from oracle import lucaskanade as ora
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.motion.lucaskanade import dense_lucaskanade as lk
    rng = np.random.default_rng(9)
    for it in range(12):
        m, n, T = int(rng.integers(60, 260)), int(rng.integers(60, 260)), int(rng.choice([2, 3, 3]))
        fr = syn.rain_frames(m, n, T, int(rng.integers(0, 1000)), dx=int(rng.integers(-3, 4)), dy=int(rng.integers(-3, 4)))
        kw = dict(dense=False, k_outlier=int(rng.choice([5, 30, 100])), nr_std_outlier=float(rng.choice([1, 2, 3])))
        xy, uv = lk(fr, **kw)
        with ora.knn_mode("ckdtree"):
            oxy, ouv = ora.dense_lucaskanade(fr, **kw)
        assert np.array_equal(xy, oxy) and np.array_equal(uv, ouv), (it, m, n, T, kw)


def test_exact_ties_dense_field_vs_oracle_ckdtree_mode(env):
    """Dense field within 1e-12 of the oracle in cKDTree mode at EVERY pixel (that oracle mode is
    bit-identical to the reference)."""
    from oracle import lucaskanade as ora
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.motion.lucaskanade import dense_lucaskanade as lk
    for seed, (m, n, T) in enumerate([(120, 160, 2), (200, 176, 3), (257, 300, 3)]):
        fr = syn.rain_frames(m, n, T, seed, dx=2, dy=-1)
        V = lk(fr)
        with ora.knn_mode("ckdtree"):
            Vo = ora.dense_lucaskanade(fr)
        assert V.shape == Vo.shape and np.abs(V - Vo).max() <= 1e-12, (m, n, T)


def test_float32_frames_vs_oracle(env):
    """float32 frames are scaled to uint8 in float32 arithmetic (B200_QUANTISE_F32), like NumPy
    does for a float32 array; sparse vectors bit-identical to the oracle, dense field <= 1e-12."""
    from oracle import lucaskanade as ora
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.motion.lucaskanade import dense_lucaskanade as lk
    rng = np.random.default_rng(0)
    for it in range(6):
        m, n = int(rng.integers(80, 260)), int(rng.integers(80, 260))
        fr = syn.rain_frames(m, n, int(rng.choice([2, 3])), it, dx=2, dy=-1)
        if it % 2:
            fr = np.where(fr > 0.1, 10 * np.log10(np.maximum(fr, 0.1)), -15.0)
        fr = (fr + 0.37 * rng.standard_normal(fr.shape)).astype(np.float32)
        xy, uv = lk(fr, dense=False)
        oxy, ouv = ora.dense_lucaskanade(fr, dense=False)
        assert np.array_equal(xy, oxy) and np.array_equal(uv, ouv), it
        V, Vo = lk(fr), ora.dense_lucaskanade(fr)
        assert V.dtype == np.float64 and np.abs(V - Vo).max() <= 1e-12, it


def test_finiteness_certificate_dies_with_a_write(env):
    """The dense LK field is finite by construction, so the extrapolator skips its finiteness scan of it --
    but only while the field is the one LK produced: any in-place write (device tensor: version counter;
    NumPy result: content fingerprint) brings the scan, and the reference's ValueError, back."""
    torch, _ = env
    import pysteps_b200
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.extrapolation import semilagrangian as slmod
    lk = pysteps_b200.motion.get_method("lk")
    extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
    frames = syn.rain_frames(192, 224, 2, 1)
    P = frames[-1]
    # device tensors
    V = lk(torch.from_numpy(frames).cuda())
    assert slmod._known_finite(V)
    ref = extrap(torch.from_numpy(P).cuda(), V, 2)
    assert torch.equal(torch.nan_to_num(extrap(torch.from_numpy(P).cuda(), V.clone(), 2), nan=-1.0),
                       torch.nan_to_num(ref, nan=-1.0)), "certified and scanned calls agree"
    V[1, 5, 7] = float("nan")
    assert not slmod._known_finite(V)
    with pytest.raises(ValueError, match="velocity contains non-finite"):
        extrap(torch.from_numpy(P).cuda(), V, 2)
    # NumPy arrays
    Vh = lk(frames)
    assert np.isfinite(Vh).all()
    extrap(P, Vh, 2)
    Vh[0, 0, 0] = np.inf
    with pytest.raises(ValueError, match="velocity contains non-finite"):
        extrap(P, Vh, 2)
