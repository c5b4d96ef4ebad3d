"""CPU check of the ERROR BOUND behind the opt-in float32-tap trajectory kernel (csrc/sl.cu, sl_f32_kernel).

The kernel certifies that its float32-sampled trajectory floors to the same tap indices as the exact
float64 one from a per-pixel bound E >= |displacement - exact displacement|.  The kernel's OUTPUTS are
checked on the GPU (tests/test_sl_gpu.py); this module checks the MATHEMATICS: a NumPy restatement of the
kernel's per-pixel recurrences (float32 lerps, the same slope / rounding coefficients, the same
certification rule) is run beside the oracle's exact trajectories on smooth AND deliberately rough
advection fields, and wherever the restatement certifies a pixel the bound must dominate the observed
deviation and the floors must agree -- at every leadtime."""
import numpy as np
import pytest

from oracle import semilagrangian as ora
from pysteps_b200 import _synthetic as syn

F = np.float32
EPS = F(1.1920929e-07)
INFLATE = F(1.001)


def _fmaf(a, b, c):
    """float32 fused multiply-add (the float64 product of two float32 values is exact)"""
    return (a.astype(np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


def _sample(Vf, cy, cx, e):
    """sample_f32 of csrc/sl.cu for all pixels: (ok, vx, vy, err_coef, iy, ix)"""
    m, n = Vf.shape[1:]
    fy, fx = np.floor(cy), np.floor(cx)
    interior = (fy >= 0) & (fx >= 0) & (fy <= m - 2) & (fx <= n - 2)
    ty, tx = (cy - fy).astype(F), (cx - fx).astype(F)
    ok = interior & (np.maximum(np.abs(ty - F(0.5)), np.abs(tx - F(0.5))) < F(0.5) - (e + F(2) * EPS))
    iy = np.where(interior, fy, 0).astype(np.int64)
    ix = np.where(interior, fx, 0).astype(np.int64)
    out = []
    g = np.zeros(cy.shape, F)
    for c in range(2):
        a00, a01, a10, a11 = Vf[c][iy, ix], Vf[c][iy, ix + 1], Vf[c][iy + 1, ix], Vf[c][iy + 1, ix + 1]
        d0, d1 = (a01 - a00).astype(F), (a11 - a10).astype(F)
        x0, x1 = _fmaf(tx, d0, a00), _fmaf(tx, d1, a10)
        w = (x1 - x0).astype(F)
        out.append(_fmaf(ty, w, x0))
        g = np.maximum(g, _fmaf(F(2), (np.abs(d0) + np.abs(d1)).astype(F), np.abs(w)))
    av = np.maximum(np.abs(out[0]), np.abs(out[1]))
    ec = _fmaf(g, e, EPS * _fmaf(F(4), g, F(4.5) * av))
    return ok, out[0], out[1], ec, np.where(interior, fy, np.nan), np.where(interior, fx, np.nan)


def _float32_trajectories(V, T):
    """displacement, bound and certificate after every leadtime (n_iter = 1, unit timesteps, fresh start)"""
    m, n = V.shape[1:]
    Vf = V.astype(F)
    gy, gx = np.meshgrid(np.arange(m, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
    dx, dy = np.zeros((m, n)), np.zeros((m, n))
    ux, uy = Vf[0].copy(), Vf[1].copy()            # s0 = 1
    E = np.zeros((m, n), F)
    Eu = F(2) * EPS * np.maximum(np.abs(ux), np.abs(uy))
    good = np.ones((m, n), bool)
    sa = INFLATE
    res = []
    for _ in range(T):
        hx, hy = dx - 0.5 * ux.astype(np.float64), dy - 0.5 * uy.astype(np.float64)
        ok, vx, vy, ec, _, _ = _sample(Vf, gy + hy, gx + hx, _fmaf(F(0.5), Eu, E))
        good &= ok
        ux, uy = vx, vy
        dx, dy = dx - ux.astype(np.float64), dy - uy.astype(np.float64)
        E = ((E + sa * ec).astype(F) * INFLATE + F(1e-12)).astype(F)
        ok, vx, vy, ec, fy, fx = _sample(Vf, gy + dy, gx + dx, E)
        good &= ok
        ux, uy = vx, vy
        Eu = (sa * ec).astype(F)
        res.append((dx.copy(), dy.copy(), E.copy(), good.copy(), fy, fx))
    return res


def _fields():
    m, n = 72, 88
    rng = np.random.default_rng(5)
    yield "smooth", syn.velocity_field(m, n, 1, "smooth") + np.array([0.37, 0.21]).reshape(2, 1, 1)
    yield "rotation", syn.velocity_field(m, n, 1, "rotation") * 8.0 + 0.123
    yield "white noise, |V| ~ 1", rng.normal(size=(2, m, n)) + 0.4
    yield "white noise, |V| ~ 6 (slopes of several px per px)", rng.normal(size=(2, m, n)) * 6.0
    steps = np.where(rng.random((2, m, n)) < 0.5, -2.3, 3.1) + 1e-3 * rng.normal(size=(2, m, n))
    yield "two-valued field (discontinuities everywhere)", steps
    yield "float32-exact integers", np.round(rng.normal(size=(2, m, n)) * 3.0)


@pytest.mark.parametrize("name,V", list(_fields()), ids=[f[0] for f in _fields()])
def test_bound_dominates_the_deviation_wherever_a_pixel_is_certified(name, V):
    m, n = V.shape[1:]
    P = syn.rain_field(m, n, 2)
    gy, gx = np.meshgrid(np.arange(m, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
    T = 8
    fast = _float32_trajectories(V, T)
    certified_any = 0
    for k in range(T):
        _, dex = ora.extrapolate(P, V, k + 1, return_displacement=True)
        dx, dy, E, good, fy, fx = fast[k]
        certified_any += int(good.sum())
        if not good.any():
            continue
        dev = np.maximum(np.abs(dx - dex[0]), np.abs(dy - dex[1]))[good]
        bound = E.astype(np.float64)[good]
        assert (dev <= bound).all(), (f"{name}, leadtime {k + 1}: deviation {dev.max():.3e} exceeds the bound "
                                      f"at {int((dev > bound).sum())} certified pixels")
        assert np.array_equal(np.floor(gx + dex[0])[good], fx[good]), f"{name}, leadtime {k + 1}: column indices"
        assert np.array_equal(np.floor(gy + dex[1])[good], fy[good]), f"{name}, leadtime {k + 1}: row indices"
    if name in ("smooth", "rotation"):
        assert certified_any > 0.5 * T * m * n  # the rule is not vacuous: most pixels ARE certified


def test_bound_stays_small_on_a_smooth_field():
    """The share of uncertified pixels is what the fix-up launch costs.  The bound grows geometrically with
    the cell slopes (exp of the summed Lipschitz constants -- here up to 0.2 px per px -- and the slope
    estimate 2(|d0|+|d1|)+|w| is up to ~2x the true one): after 12 leadtimes it must still be far below the
    cell size -- 1e-4 px in the median, 1e-3 at worst (observed deviations on the GPU: ~2e-6)."""
    V = syn.velocity_field(96, 112, 3, "smooth") + np.array([0.37, 0.21]).reshape(2, 1, 1)
    _, _, E, good, _, _ = _float32_trajectories(V, 12)[-1]
    assert float(np.median(E[good])) < 3e-4 and float(E[good].max()) < 1e-3
