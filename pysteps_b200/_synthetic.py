"""Seeded synthetic radar frames and advection fields.

NumPy only; used by tests/, bench.py and __graft_entry__.smoke() so that the CPU
oracle and the CUDA path see identical inputs.  All values are exactly
representable in float32 (generated in float64, rounded through float32)."""
import numpy as np


def _f32_exact(a):
    return a.astype(np.float32).astype(np.float64)


def powerlaw_field(m, n, seed=0, beta=1.5):
    """Isotropic power-law random field, standardised."""
    rng = np.random.default_rng(seed)
    w = rng.standard_normal((m, n))
    ky = np.fft.fftfreq(m)[:, None] * m
    kx = np.fft.rfftfreq(n)[None, :] * n
    k = np.sqrt(ky * ky + kx * kx)
    k[0, 0] = 1.0
    f = np.fft.irfft2(np.fft.rfft2(w) * k ** (-beta), s=(m, n))
    return (f - f.mean()) / f.std()


def rain_field(m, n, seed=0):
    """~30 % wet area, rain rates 5-40, zeros elsewhere (float64, f32-exact)."""
    f = powerlaw_field(m, n, seed)
    return _f32_exact(np.where(f > 0.5, 10.0 * f, 0.0))


def shift_frame(r, dy, dx):
    """Integer translation with zero inflow (frame(t+1)[y,x] = frame(t)[y-dy, x-dx])."""
    out = np.zeros_like(r)
    m, n = r.shape
    ys, yd = (slice(0, m - dy), slice(dy, m)) if dy >= 0 else (slice(-dy, m), slice(0, m + dy))
    xs, xd = (slice(0, n - dx), slice(dx, n)) if dx >= 0 else (slice(-dx, n), slice(0, n + dx))
    out[yd, xd] = r[ys, xs]
    return out


def rain_frames(m, n, nframes=2, seed=0, dx=3, dy=-2):
    """(nframes, m, n) sequence translating by (dx, dy) px per step."""
    r0 = rain_field(m, n, seed)
    return np.stack([shift_frame(r0, dy * k, dx * k) for k in range(nframes)])


def _box9(a):
    pad = np.pad(a, 4, mode="edge")
    c = np.cumsum(np.cumsum(pad, axis=0), axis=1)
    c = np.pad(c, ((1, 0), (1, 0)))
    m, n = a.shape
    return (c[9:9 + m, 9:9 + n] - c[0:m, 9:9 + n] - c[9:9 + m, 0:n] + c[0:m, 0:n]) / 81.0


def velocity_field(m, n, seed=0, kind="smooth", u=3.0, v=-2.0):
    """(2, m, n) advection field in px/step; [0] = x component, [1] = y component.
    kind: "smooth" (mean (u,v) + smoothed noise), "rotation" (solid body, 0.002 rad/step),
    "uniform"."""
    rng = np.random.default_rng(seed + 1000)
    if kind == "uniform":
        V = np.stack([np.full((m, n), u), np.full((m, n), v)])
    elif kind == "smooth":
        V = np.stack([u + _box9(0.1 * rng.standard_normal((m, n)) * 9.0),
                      v + _box9(0.1 * rng.standard_normal((m, n)) * 9.0)])
    elif kind == "rotation":
        y, x = np.mgrid[0:m, 0:n].astype(np.float64)
        w = 0.002
        V = np.stack([-w * (y - (m - 1) / 2.0), w * (x - (n - 1) / 2.0)]) * 20.0
    else:
        raise ValueError(kind)
    return _f32_exact(V)


def nan_disc(r, frac=0.125):
    """Copy of r with a centred disc of NaN of radius frac*m."""
    m, n = r.shape
    y, x = np.mgrid[0:m, 0:n]
    out = r.copy()
    out[(y - m / 2) ** 2 + (x - n / 2) ** 2 < (frac * m) ** 2] = np.nan
    return out
