"""Oracle mirror of ``pysteps.motion.proesmans.proesmans`` -- TEST INFRASTRUCTURE ONLY
(see ``oracle/__init__.py``).  Host side restates pysteps/motion/proesmans.py:20-94 and the
driver of pysteps/motion/_proesmans.pyx:19-44,60-76; the array arithmetic runs in
``proesmans_oracle.c``.  Pinned (to a tolerance: the reference extension is built with
``-ffast-math``) by tests/test_oracle_proesmans.py against the reference extension compiled out
of tree with the reference's flags."""
import ctypes

import numpy as np

from . import lib

_dp = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64


def _p(a):
    return a.ctypes.data_as(_dp)


def raster_order_mean(on):
    """Accumulate the mean inconsistency in the literal raster order of _proesmans.pyx:209-228
    (pinning against an IEEE build of the reference source) instead of row by row (default)."""
    lib().ora_proesmans_raster_sum(int(bool(on)))


def construct_image_pyramid(R, n_levels):
    """_proesmans.pyx:60-76"""
    L = lib()
    L.ora_proesmans_pyr_down.restype = None
    out = [np.ascontiguousarray(R, dtype=np.float64)]
    m, n = out[0].shape
    for _ in range(1, n_levels):
        nxt = np.zeros((int(m / 2), int(n / 2)))
        L.ora_proesmans_pyr_down(_p(out[-1]), _i64(out[-1].shape[0]), _i64(out[-1].shape[1]), _p(nxt))
        out.append(nxt)
        m, n = int(m / 2), int(n / 2)
    return out


def consistency_maps(V):
    """_proesmans.pyx:190-254"""
    V = np.ascontiguousarray(V, dtype=np.float64)
    G = np.empty((2,) + V.shape[2:])
    L = lib()
    L.ora_proesmans_consistency.restype = None
    L.ora_proesmans_consistency(_p(V), _i64(V.shape[2]), _i64(V.shape[3]), _p(G))
    return G


def compute_advection_field(R, lam, num_iter, n_levels):
    """_proesmans.pyx:19-44"""
    L = lib()
    L.ora_proesmans_level.restype = ctypes.c_int
    L.ora_proesmans_next_level.restype = None
    R = np.ascontiguousarray(R, dtype=np.float64)
    pyr = [construct_image_pyramid(R[0], n_levels), construct_image_pyramid(R[1], n_levels)]
    m, n = pyr[0][-1].shape
    V = np.zeros((2, 2, m, n))
    for i in range(n_levels - 1, -1, -1):
        Ri = np.ascontiguousarray(np.stack([pyr[0][i], pyr[1][i]]))
        h, w = Ri.shape[1:]
        rc = L.ora_proesmans_level(_p(Ri), _i64(h), _i64(w), _p(V), _i64(int(num_iter)), ctypes.c_double(lam))
        if rc != 0:
            raise MemoryError("oracle allocation failed")
        if i > 0:
            hn, wn = pyr[0][i - 1].shape
            Vn = np.zeros((2, 2, hn, wn))
            L.ora_proesmans_next_level(_p(V), _i64(h), _i64(w), _p(Vn), _i64(hn), _i64(wn))
            V = Vn
    return V, consistency_maps(V)


def proesmans(input_images, lam=50.0, num_iter=100, num_levels=6, filter_std=0.0, verbose=True,
              full_output=False):
    """pysteps/motion/proesmans.py:20-94 behind decorators.check_input_frames(2, 2)."""
    if input_images.ndim != 3:
        raise ValueError("input_images dimension mismatch.\n"
                         f"input_images.shape: {str(input_images.shape)}\n"
                         "(t, x, y ) dimensions expected")
    if 2 < input_images.shape[0] > 2:
        raise ValueError(f"input_images frames {input_images.shape[0]} mismatch.\n"
                         "Minimum frames: 2\nMaximum frames: 2\n")
    del verbose
    im = np.stack([input_images[-2, :, :].copy(), input_images[-1, :, :].copy()])
    im_min, im_max = np.min(im), np.max(im)
    if im_max - im_min > 1e-8:
        im = (im - im_min) / (im_max - im_min) * 255.0
    if filter_std > 0.0:  # proesmans.py:85-87 (scipy itself: the reference's own dependency)
        from scipy.ndimage import gaussian_filter
        im[0, :, :] = gaussian_filter(im[0, :, :], filter_std)
        im[1, :, :] = gaussian_filter(im[1, :, :], filter_std)
    advfield, quality = compute_advection_field(im, lam, num_iter, num_levels)
    if not full_output:
        return advfield[0]
    return advfield, quality
