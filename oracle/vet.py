"""Oracle mirror of ``pysteps.motion.vet`` -- TEST INFRASTRUCTURE ONLY.

The native extension of the reference (pysteps/motion/_vet.pyx) is restated in
``vet_oracle.c`` (pinned to the reference's own compiled extension at relative 1e-9,
tests/golden/gen_vet_golden.py); the driver below restates pysteps/motion/vet.py:165-648.
The optimiser is SciPy's ``minimize`` as in the reference (third-party, unpinned:
requirements.txt:5); ``scipy.ndimage.zoom`` is restated in C (bit-exact, tests).
"""
import ctypes

import numpy
from numpy.ma.core import MaskedArray
from scipy.optimize import minimize

from . import lib

_dp = ctypes.POINTER(ctypes.c_double)
_i8p = ctypes.POINTER(ctypes.c_int8)
_i64 = ctypes.c_int64


def zoom_o1(a, oh, ow):
    a = numpy.ascontiguousarray(a, dtype=numpy.float64)
    c, h, w = a.shape
    out = numpy.empty((c, oh, ow))
    L = lib()
    L.ora_zoom_o1.restype = None
    L.ora_zoom_o1(a.ctypes.data_as(_dp), _i64(c), _i64(h), _i64(w), _i64(oh), _i64(ow),
                  out.ctypes.data_as(_dp))
    return out


def warp(image, mask, displacement, gradient=False):
    """_vet._warp (pysteps/motion/_vet.pyx:66-232)."""
    image = numpy.ascontiguousarray(image, dtype=numpy.float64)
    mask = numpy.ascontiguousarray(mask, dtype=numpy.int8)
    displacement = numpy.ascontiguousarray(displacement, dtype=numpy.float64)
    nx, ny = image.shape
    out = numpy.empty((nx, ny))
    omask = numpy.empty((nx, ny), numpy.int8)
    grad = numpy.empty((2, nx, ny)) if gradient else None
    L = lib()
    L.ora_vet_warp.restype = None
    L.ora_vet_warp(image.ctypes.data_as(_dp), mask.ctypes.data_as(_i8p), displacement.ctypes.data_as(_dp),
                   _i64(nx), _i64(ny), out.ctypes.data_as(_dp), omask.ctypes.data_as(_i8p),
                   None if grad is None else grad.ctypes.data_as(_dp))
    return (out, omask, grad) if gradient else (out, omask)


def cost_function(sector_displacement, template_image, input_image, mask, smooth_gain, gradient=False):
    """_vet._cost_function (pysteps/motion/_vet.pyx:238-621)."""
    sd = numpy.ascontiguousarray(sector_displacement, dtype=numpy.float64)
    t = numpy.ascontiguousarray(template_image, dtype=numpy.float64)
    inp = numpy.ascontiguousarray(input_image, dtype=numpy.float64)
    mk = numpy.ascontiguousarray(mask, dtype=numpy.int8)
    xs, ys = sd.shape[1:]
    nx, ny = t.shape
    out = numpy.empty((2, xs, ys) if gradient else 2)
    L = lib()
    L.ora_vet_cost.restype = ctypes.c_int
    rc = L.ora_vet_cost(sd.ctypes.data_as(_dp), t.ctypes.data_as(_dp), inp.ctypes.data_as(_dp),
                        mk.ctypes.data_as(_i8p), _i64(xs), _i64(ys), _i64(nx), _i64(ny),
                        ctypes.c_float(smooth_gain), ctypes.c_int(1 if gradient else 0),
                        out.ctypes.data_as(_dp))
    if rc == -1:
        raise ValueError("Error computing cost function.\n", "The number of sectors don't divide the image size")
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return out if gradient else (out[0], out[1])


def vet_cost_function_gradient(*args, **kwargs):
    kwargs["gradient"] = True
    return vet_cost_function(*args, **kwargs)


def vet_cost_function(sector_displacement_1d, input_images, blocks_shape, mask, smooth_gain,
                      debug=False, gradient=False):
    """pysteps/motion/vet.py:165-299."""
    sd = sector_displacement_1d.reshape(*((2,) + tuple(blocks_shape)))
    if input_images.shape[0] == 3:
        three_times = True
        previous_image, center_image, next_image = input_images[0], input_images[1], input_images[2]
    else:
        three_times = False
        previous_image, center_image, next_image = None, input_images[0], input_images[1]
    if gradient:
        g = cost_function(sd, center_image, next_image, mask, smooth_gain, gradient=True)
        if three_times:
            g += cost_function(sd, previous_image, center_image, mask, smooth_gain, gradient=True)
        return g.ravel()
    residuals, smoothness_penalty = cost_function(sd, center_image, next_image, mask, smooth_gain)
    if three_times:
        _r, _s = cost_function(sd, previous_image, center_image, mask, smooth_gain)
        residuals += _r
        smoothness_penalty += _s
    return residuals + smoothness_penalty


def get_padding(dimension_size, sectors):
    reminder = dimension_size % sectors
    if reminder != 0:
        pad = sectors - reminder
        pad_before = pad // 2
        pad_after = pad_before if pad % 2 == 0 else pad_before + 1
        return pad_before, pad_after
    return 0, 0


def vet(input_images, sectors=((32, 16, 4, 2), (32, 16, 4, 2)), smooth_gain=1e6, first_guess=None,
        intermediate_steps=False, verbose=False, indexing="yx", padding=0, options=None):
    """pysteps/motion/vet.py:302-648."""
    # decorators.check_input_frames(2, 3), pysteps/decorators.py:121-146
    if input_images.ndim != 3:
        raise ValueError("input_images dimension mismatch.\n"
                         f"input_images.shape: {str(input_images.shape)}\n"
                         "(t, x, y ) dimensions expected")
    if 2 < input_images.shape[0] > 3:
        raise ValueError(f"input_images frames {input_images.shape[0]} mismatch.\n"
                         "Minimum frames: 2\nMaximum frames: 3\n")
    options = dict() if options is None else dict(options)
    options.setdefault("eps", 0.1)
    options.setdefault("gtol", 0.1)
    options.setdefault("maxiter", 100)
    options.setdefault("disp", False)
    method = options.pop("method", "CG")
    if indexing not in ["yx", "xy", "ij"]:
        raise ValueError("Invalid indexing values: {0}\n".format(indexing)
                         + "Supported values: {0}".format(str(["yx", "xy", "ij"])))
    if not isinstance(input_images, MaskedArray):
        input_images = numpy.ma.masked_invalid(input_images)
    else:
        input_images = input_images.copy()
    mask = numpy.ma.getmaskarray(input_images)
    if padding > 0:
        pt = ((0, 0), (padding, padding), (padding, padding))
        data = numpy.pad(numpy.ma.getdata(input_images), pt, "constant", constant_values=numpy.nan)
        mask = numpy.pad(mask, pt, "constant", constant_values=True)
        input_images = numpy.ma.MaskedArray(data=data, mask=mask)
    input_images.data[mask] = 0
    mask = numpy.asarray(numpy.any(mask, axis=0), dtype="int8", order="C")
    input_images = numpy.asarray(input_images.data, dtype="float64", order="C")
    sectors = numpy.asarray(sectors, dtype="int", order="C")
    if sectors.ndim == 1:
        sectors = numpy.zeros((2,) + sectors.shape, dtype="int", order="C") + sectors.reshape(
            (1, sectors.shape[0]))
    elif sectors.ndim > 2 or sectors.ndim < 1:  # vet.py:513-519
        raise ValueError("Incorrect sectors dimensions.\n"
                         + "Only 1D or 2D arrays are supported to define"
                         + "the number of sectors used in"
                         + "the scaling procedure")
    sectors[0, :].sort()
    sectors[1, :].sort()
    fgs = (2, int(sectors[0, 0]), int(sectors[1, 0]))
    if first_guess is None:
        first_guess = numpy.zeros(fgs, order="C")
    elif first_guess.shape != fgs:  # vet.py:531-537
        raise ValueError("The shape of the initial guess do not match the number of "
                         + "sectors of the first scaling guess\n"
                         + "first_guess.shape={}\n".format(str(first_guess.shape))
                         + "Expected shape={}".format(str(fgs)))
    else:
        first_guess = numpy.asarray(first_guess, order="C", dtype="float64")
    scaling_guesses = []
    psi, psj = sectors[0, 0], sectors[1, 0]
    for n, (si, sj) in enumerate(zip(sectors[0, :], sectors[1, :])):
        pad_i = get_padding(input_images.shape[1], si)
        pad_j = get_padding(input_images.shape[2], sj)
        if (pad_i != (0, 0)) or (pad_j != (0, 0)):
            _images = numpy.pad(input_images, ((0, 0), pad_i, pad_j), "edge")
            _mask = numpy.ascontiguousarray(numpy.pad(mask, (pad_i, pad_j), "constant", constant_values=1))
        else:
            _images, _mask = input_images, mask
        if n > 0:
            first_guess = zoom_o1(first_guess, int(round(first_guess.shape[1] * (si / psi))),
                                  int(round(first_guess.shape[2] * (sj / psj))))
        result = minimize(vet_cost_function, first_guess.flatten(), jac=vet_cost_function_gradient,
                          args=(_images, (si, sj), _mask, smooth_gain), method=method, options=options)
        first_guess = result.x.reshape(*first_guess.shape)
        scaling_guesses.append(first_guess[::-1, ...] if indexing == "yx" else first_guess)
        psi, psj = si, sj
    ni, nj = _images.shape[1], _images.shape[2]
    first_guess = zoom_o1(first_guess, int(round(first_guess.shape[1] * (ni / si))),
                          int(round(first_guess.shape[2] * (nj / sj))))
    first_guess = first_guess[:, pad_i[0]: ni - pad_i[1], pad_j[0]: nj - pad_j[1]]
    if indexing == "yx":
        first_guess = first_guess[::-1, ...]
    if padding > 0:
        first_guess = first_guess[:, padding:-padding, padding:-padding]
    if intermediate_steps:
        return first_guess, scaling_guesses
    return first_guess
