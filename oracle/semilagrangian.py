"""Oracle mirror of ``pysteps.extrapolation.semilagrangian.extrapolate``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Host-side argument
handling restates pysteps/extrapolation/semilagrangian.py:106-179,260-266 line
by line; the array arithmetic runs in ``sl_oracle.c`` / ``spline_oracle.c``.
Every ``interp_order`` scipy accepts (0 .. 5) is restated.
"""
import ctypes
import warnings

import numpy as np

from . import lib

_MODES = {"constant": 0, "nearest": 1}
_dp = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def spline_filter(a, order, mode="constant"):
    """scipy.ndimage.spline_filter(a, order, output=float64, mode=mode) for the two boundary
    treatments map_coordinates uses: "constant" (mirror) and "nearest" (reflect)."""
    f = np.array(a, dtype=np.float64, order="C")
    L = lib()
    L.ora_spline_filter.restype = None
    L.ora_spline_filter.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    L.ora_spline_filter(_p(f), f.shape[0], f.shape[1], int(order), int(_MODES[mode] == 1))
    return f


def spline_filter3(a, mode="constant"):
    return spline_filter(a, 3, mode)


def map_coordinates_spline(a, coords, order, mode="constant", cval=0.0):
    """scipy.ndimage.map_coordinates(a, coords, order=0|2|3|4|5, mode, cval, prefilter=True)."""
    a64 = np.ascontiguousarray(a, dtype=np.float64)
    cy = np.ascontiguousarray(coords[0], dtype=np.float64)
    cx = np.ascontiguousarray(coords[1], dtype=np.float64)
    out = np.empty(cy.shape, dtype=np.float64)
    L = lib()
    L.ora_map_coordinates_spline.restype = ctypes.c_int
    L.ora_map_coordinates_spline.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, _dp, _dp, ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp]
    rc = L.ora_map_coordinates_spline(_p(a64), a64.shape[0], a64.shape[1], _p(cy), _p(cx), cy.size,
                                      int(order), _MODES[mode], float(cval), _p(out))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return out.astype(a.dtype) if a.dtype == np.float32 else out


def map_coordinates_o1(a, coords, mode="constant", cval=0.0):
    """scipy.ndimage.map_coordinates(a, coords, order=1, prefilter=False)."""
    a64 = np.ascontiguousarray(a, dtype=np.float64)
    cy = np.ascontiguousarray(coords[0], dtype=np.float64)
    cx = np.ascontiguousarray(coords[1], dtype=np.float64)
    out = np.empty(cy.shape, dtype=np.float64)
    L = lib()
    L.ora_map_coordinates_o1.restype = None
    L.ora_map_coordinates_o1.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, _dp, _dp,
                                         ctypes.c_int64, ctypes.c_int, ctypes.c_double, _dp]
    L.ora_map_coordinates_o1(_p(a64), a64.shape[0], a64.shape[1], _p(cy), _p(cx),
                             cy.size, _MODES[mode], float(cval), _p(out))
    return out.astype(a.dtype) if a.dtype == np.float32 else out


def extrapolate(precip, velocity, timesteps, outval=np.nan, xy_coords=None,
                allow_nonfinite_values=False, vel_timestep=1, **kwargs):
    # semilagrangian.py:106-126
    if precip is not None and precip.ndim != 2:
        raise ValueError("precip must be a two-dimensional array")
    if velocity.ndim != 3:
        raise ValueError("velocity must be a three-dimensional array")
    if not allow_nonfinite_values:
        if precip is not None and np.any(~np.isfinite(precip)):
            raise ValueError("precip contains non-finite values")
        if np.any(~np.isfinite(velocity)):
            raise ValueError("velocity contains non-finite values")
    if precip is not None and np.all(~np.isfinite(precip)):
        raise ValueError("precip contains only non-finite values")
    if np.all(~np.isfinite(velocity)):
        raise ValueError("velocity contains only non-finite values")
    if isinstance(timesteps, list) and not sorted(timesteps) == timesteps:
        raise ValueError("timesteps is not in ascending order")
    # :129-134
    displacement_prev = kwargs.get("displacement_prev", None)
    n_iter = kwargs.get("n_iter", 1)
    return_displacement = kwargs.get("return_displacement", False)
    interp_order = kwargs.get("interp_order", 1)
    map_coordinates_mode = kwargs.get("map_coordinates_mode", "constant")
    if precip is None and not return_displacement:
        raise ValueError("precip is None but return_displacement is False")
    if "D_prev" in kwargs.keys():
        warnings.warn("deprecated argument D_prev is ignored, use displacement_prev instead")
    if interp_order not in (0, 1, 2, 3, 4, 5):
        raise RuntimeError("spline order not supported")  # scipy.ndimage._ni_support
    # :144-157 separate masks preserve NaN and no-precipitation values under the spline
    mask_min = mask_finite = None
    minval = 0.0
    if precip is not None and interp_order > 1:
        minval = np.nanmin(precip)
        mask_min = (precip > minval).astype(float)
        if allow_nonfinite_values:
            mask_finite = np.isfinite(precip)
            precip = precip.copy()
            precip[~mask_finite] = 0.0
            mask_finite = mask_finite.astype(float)
        else:
            mask_finite = np.ones(precip.shape)
    # :159-165
    if isinstance(timesteps, int):
        timesteps = np.arange(1, timesteps + 1)
        vel_timestep = 1.0
    elif np.any(np.diff(timesteps) <= 0.0):
        raise ValueError("the given timestep sequence is not monotonously increasing")
    timestep_diff = np.ascontiguousarray(
        np.hstack([[timesteps[0]], np.diff(timesteps)]), dtype=np.float64)
    # :171-172
    if precip is not None and isinstance(outval, str) and outval == "min":
        outval = np.nanmin(precip)
    m, n = velocity.shape[1], velocity.shape[2]
    V = np.ascontiguousarray(velocity, dtype=np.float64)
    P = None if precip is None else np.ascontiguousarray(precip, dtype=np.float64)
    XY = None if xy_coords is None else np.ascontiguousarray(xy_coords, dtype=np.float64)
    DP = None if displacement_prev is None else np.ascontiguousarray(
        displacement_prev, dtype=np.float64)
    T = timestep_diff.size
    out = None if precip is None else np.empty((T, m, n), dtype=np.float64)
    disp = np.empty((2, m, n), dtype=np.float64)
    MM = None if mask_min is None else np.ascontiguousarray(mask_min, dtype=np.float64)
    MF = None if mask_finite is None else np.ascontiguousarray(mask_finite, dtype=np.float64)
    L = lib()
    L.ora_sl_extrapolate_order.restype = ctypes.c_int
    L.ora_sl_extrapolate_order.argtypes = [_dp, _dp, ctypes.c_int64, ctypes.c_int64, _dp, _dp,
                                           ctypes.c_int64, ctypes.c_double, ctypes.c_int, _dp,
                                           ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           _dp, _dp, ctypes.c_double, _dp, _dp]
    rc = L.ora_sl_extrapolate_order(_p(P), _p(V), m, n, _p(XY), _p(timestep_diff), T,
                                    float(vel_timestep), int(n_iter), _p(DP),
                                    float(outval) if precip is not None else 0.0,  # cval unused without precip
                                    _MODES[map_coordinates_mode],
                                    int(velocity.dtype == np.float32), int(interp_order), _p(MM), _p(MF),
                                    float(minval), _p(out), _p(disp))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    # :260-266 ; scipy returns the input array dtype
    if precip is not None:
        if precip.dtype != np.float64:
            out = out.astype(precip.dtype)
        if not return_displacement:
            return out
        return out, disp
    return None, disp
