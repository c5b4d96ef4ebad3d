"""B200 Variational Echo Tracking -- drop-in for ``pysteps.motion.vet``
(pysteps/motion/vet.py:93-648; native extension pysteps/motion/_vet.pyx -> csrc/vet.cu).

Design (not the reference's control flow):

* The raw frames are uploaded ONCE.  Every minimisation level derives its image stack and mask
  from them on the device in one kernel (``b200_vet_level_images``: NaN / user-mask cleaning,
  the global ``padding`` ring and the level's divisibility padding) -- nothing of image size is
  built on the host, and a level whose padding equals an earlier one reuses that stack.
* The objective is evaluated as a PAIR: value and gradient at the same point from one pass over
  the images (``b200_vet_value_and_gradient``: one kernel per frame pair + two tiny finalisers,
  one 16 KB upload, one read-back).  A line search asks for both at every trial step, so the
  ~800 separate cost / gradient evaluations of a 2048^2 field become ~400 fused ones.
  ``_Objective`` memoises the last point; SciPy's CG (the optimiser the reference uses,
  vet.py:593-600 -- same iterates, same stopping rule) sees an ordinary ``fun`` / ``jac``.
* Sector fields move between levels through the device zoom (``b200_zoom_bilinear``,
  bit-identical to ``scipy.ndimage.zoom(order=1, mode="nearest")``), the final field is zoomed,
  cropped and flipped on the device and crosses PCIe once.
"""
import numpy
import torch
from numpy.ma.core import MaskedArray
from scipy.optimize import minimize

from .. import _device, _lib

_INDEXING = ("yx", "xy", "ij")


def round_int(scalar):
    return int(numpy.round(scalar))


def ceil_int(scalar):
    return int(numpy.ceil(scalar))


def get_padding(dimension_size, sectors):
    """(before, after) so that `sectors` divides the padded size; the odd cell goes after
    (vet.py:55-90)."""
    missing = (-int(dimension_size)) % int(sectors)
    return missing // 2, missing - missing // 2


def _stream():
    return _device.stream_ptr()


def morph(image, displacement, gradient=False):
    """Morph an image by a displacement field (vet.py:93-153 -> _vet._warp).
    Returns (image, mask) or (image, mask, gradient) as NumPy arrays."""
    _device.require_cuda()
    masked = isinstance(image, MaskedArray)
    d_mask = _device.to_device(numpy.ascontiguousarray(numpy.ma.getmaskarray(image), dtype=numpy.int8)) if masked \
        else torch.zeros(tuple(numpy.shape(image)), dtype=torch.int8, device="cuda")
    d_img = _device.to_device(numpy.ascontiguousarray(numpy.ma.getdata(image), dtype=numpy.float64))
    d_disp = _device.to_device(numpy.ascontiguousarray(displacement, dtype=numpy.float64))
    nx, ny = (int(v) for v in d_img.shape)
    out = torch.empty((nx, ny), dtype=torch.float64, device="cuda")
    omask = torch.empty((nx, ny), dtype=torch.int8, device="cuda")
    grad = torch.empty((2, nx, ny), dtype=torch.float64, device="cuda") if gradient else None
    _lib.call("b200_vet_warp", d_img.data_ptr(), d_mask.data_ptr(), d_disp.data_ptr(), nx, ny,
              out.data_ptr(), omask.data_ptr(), _device.ptr(grad), _stream())
    res = (out.cpu().numpy(), omask.cpu().numpy())
    return res + (grad.cpu().numpy(),) if gradient else res


# ---------------------------------------------------------------------------------------------
class _LevelImages:
    """Image stack (T, M, N) float64 and mask (M, N) int8 of one level, resident in HBM."""

    def __init__(self, images, mask):
        self.images, self.mask = images, mask
        self.shape = tuple(int(v) for v in images.shape)

    @classmethod
    def from_host(cls, images, mask):
        return cls(_device.to_device(numpy.ascontiguousarray(images, dtype=numpy.float64)),
                   _device.to_device(numpy.ascontiguousarray(mask, dtype=numpy.int8)))


class _Frames:
    """The raw input on the device; hands out the level stacks."""

    def __init__(self, input_images, padding):
        user_mask = None
        if isinstance(input_images, MaskedArray):
            # a MaskedArray is taken at its word: only its own mask (vet.py:500-505)
            user_mask = _device.to_device(numpy.ascontiguousarray(numpy.ma.getmaskarray(input_images),
                                                                  dtype=numpy.uint8))
        self.frames = _device.to_device(numpy.ascontiguousarray(numpy.ma.getdata(input_images), dtype=numpy.float64))
        self.user_mask = user_mask
        self.T, self.m, self.n = (int(v) for v in self.frames.shape)
        self.padding = int(padding)
        self.mg, self.ng = self.m + 2 * self.padding, self.n + 2 * self.padding  # globally padded frame
        self._levels = {}

    def level(self, pad_i, pad_j):
        key = (pad_i, pad_j)
        if key not in self._levels:
            M, N = self.mg + pad_i[0] + pad_i[1], self.ng + pad_j[0] + pad_j[1]
            images = torch.empty((self.T, M, N), dtype=torch.float64, device="cuda")
            mask = torch.empty((M, N), dtype=torch.int8, device="cuda")
            _lib.call("b200_vet_level_images", self.frames.data_ptr(), _device.ptr(self.user_mask), self.T,
                      self.m, self.n, self.padding, pad_i[0], pad_j[0], M, N, images.data_ptr(),
                      mask.data_ptr(), _stream())
            self._levels[key] = _LevelImages(images, mask)
        return self._levels[key]


class _Objective:
    """VET cost and gradient on one level for a (2, bi, bj) sector field, evaluated as a pair and
    memoised on the last point."""

    def __init__(self, level, blocks_shape, smooth_gain):
        bi, bj = int(blocks_shape[0]), int(blocks_shape[1])
        if bi < 2 or bj < 2:
            raise NotImplementedError("pysteps_b200 VET: at least 2 x 2 sectors are required")
        if level.shape[1] % bi != 0 or level.shape[2] % bj != 0:
            raise ValueError("Error computing cost function.\n",
                             "The number of sectors don't divide the image size")
        self.level, self.bi, self.bj = level, bi, bj
        self.gain = float(smooth_gain)
        self.size = 2 * bi * bj
        self.work = torch.empty(3 * self.size + 4, dtype=torch.float64, device="cuda")
        self._x = None
        self._parts = numpy.zeros(2)
        self._grad = numpy.zeros(self.size)

    def _at(self, x):
        x = numpy.ascontiguousarray(x, dtype=numpy.float64).reshape(-1)
        if x.size != self.size:
            raise ValueError(f"cannot reshape array of size {x.size} into shape {(2, self.bi, self.bj)}")
        if self._x is None or not numpy.array_equal(x, self._x):
            T, nx, ny = self.level.shape
            _lib.call("b200_vet_value_and_gradient", x.ctypes.data, self.level.images.data_ptr(), T,
                      self.level.mask.data_ptr(), self.bi, self.bj, nx, ny, self.gain, self.work.data_ptr(),
                      self._parts.ctypes.data, self._grad.ctypes.data, _stream())
            self._x = x.copy()
        return self

    def parts(self, x):
        """(residuals, smoothness penalty)"""
        self._at(x)
        return float(self._parts[0]), float(self._parts[1])

    def value(self, x):
        r, s = self.parts(x)
        return r + s

    def gradient(self, x):
        return self._at(x)._grad.copy()


def vet_cost_function_gradient(*args, **kwargs):
    """Gradient of the VET cost function (vet.py:156-162)."""
    kwargs["gradient"] = True
    return vet_cost_function(*args, **kwargs)


def vet_cost_function(sector_displacement_1d, input_images, blocks_shape, mask, smooth_gain,
                      debug=False, gradient=False):
    """VET cost function / gradient for a flattened (2, bi, bj) sector field (vet.py:165-299).
    `input_images`: the (2 or 3, nx, ny) NumPy stack of the reference with its int8 `mask`."""
    _device.require_cuda()
    level = input_images if isinstance(input_images, _LevelImages) else \
        _LevelImages.from_host(numpy.asarray(input_images), numpy.asarray(mask))
    if level.shape[0] not in (2, 3):
        raise ValueError("vet_cost_function needs two or three frames")
    obj = _Objective(level, blocks_shape, smooth_gain)
    if gradient:
        return obj.gradient(sector_displacement_1d)
    residuals, smoothness_penalty = obj.parts(sector_displacement_1d)
    if debug:
        print("\nresiduals", residuals)
        print("smoothness_penalty", smoothness_penalty)
    return residuals + smoothness_penalty


# ---------------------------------------------------------------------------------------------
def _sector_table(sectors):
    """(levels, 2) array of sector counts per axis, coarse to fine (vet.py:525-543)."""
    s = numpy.asarray(sectors, dtype="int", order="C")
    if s.ndim == 1:
        s = numpy.stack([s, s])
    elif s.ndim != 2:
        raise ValueError(
            "Incorrect sectors dimensions.\n"
            + "Only 1D or 2D arrays are supported to define"
            + "the number of sectors used in"
            + "the scaling procedure"
        )
    return numpy.sort(s, axis=1).T.copy()


def _zoom_on_device(field, oh, ow):
    """scipy.ndimage.zoom(field, (1, oh/h, ow/w), order=1, mode="nearest") -> device tensor"""
    d_in = field if isinstance(field, torch.Tensor) else \
        _device.to_device(numpy.ascontiguousarray(field, dtype=numpy.float64))
    c, h, w = (int(v) for v in d_in.shape)
    out = torch.empty((c, oh, ow), dtype=torch.float64, device="cuda")
    _lib.call("b200_zoom_bilinear", d_in.data_ptr(), c, h, w, oh, ow, out.data_ptr(), _stream())
    return out


def vet(input_images, sectors=((32, 16, 4, 2), (32, 16, 4, 2)), smooth_gain=1e6, first_guess=None,
        intermediate_steps=False, verbose=True, indexing="yx", padding=0, options=None):
    """Variational Echo Tracking; same contract as the reference (vet.py:302-648)."""
    # decorators.check_input_frames(2, 3): decorators.py:121-146
    if input_images.ndim != 3:
        raise ValueError(
            "input_images dimension mismatch.\n"
            f"input_images.shape: {str(input_images.shape)}\n"
            "(t, x, y ) dimensions expected"
        )
    if 2 < input_images.shape[0] > 3:
        raise ValueError(
            f"input_images frames {input_images.shape[0]} mismatch.\n"
            "Minimum frames: 2\n"
            "Maximum frames: 3\n"
        )
    _device.require_cuda()
    say = print if verbose else (lambda *a, **k: None)

    opts = {"eps": 0.1, "gtol": 0.1, "maxiter": 100, "disp": False}
    opts.update(options or {})
    method = opts.pop("method", "CG")
    say("Running VET algorithm")
    if indexing not in _INDEXING:
        raise ValueError(
            "Invalid indexing values: {0}\n".format(indexing)
            + "Supported values: {0}".format(str(list(_INDEXING)))
        )

    frames = _Frames(input_images, padding)
    table = _sector_table(sectors)
    guess_shape = (2, int(table[0, 0]), int(table[0, 1]))
    if first_guess is None:
        x = numpy.zeros(guess_shape)
    elif first_guess.shape != guess_shape:
        raise ValueError(
            "The shape of the initial guess do not match the number of "
            + "sectors of the first scaling guess\n"
            + "first_guess.shape={}\n".format(str(first_guess.shape))
            + "Expected shape={}".format(str(guess_shape))
        )
    else:
        x = numpy.asarray(first_guess, order="C", dtype="float64")

    history = []
    level = pads = None
    for n_level, (bi, bj) in enumerate((int(a), int(b)) for a, b in table):
        pads = (get_padding(frames.mg, bi), get_padding(frames.ng, bj))
        level = frames.level(*pads)
        say(f"level {n_level}: {bi} x {bj} sectors of {level.shape[1] // bi} x {level.shape[2] // bj} px, "
            f"frame {frames.mg} x {frames.ng} padded to {level.shape[1]} x {level.shape[2]}")
        if n_level > 0:
            # the finer grid starts from the bilinear zoom of the coarser solution (vet.py:580-589)
            factor_i, factor_j = bi / x.shape[1], bj / x.shape[2]
            x = _zoom_on_device(x, int(round(x.shape[1] * factor_i)), int(round(x.shape[2] * factor_j))).cpu().numpy()
        obj = _Objective(level, (bi, bj), smooth_gain)
        result = minimize(obj.value, x.reshape(-1), jac=obj.gradient, method=method, options=opts)
        x = result.x.reshape(x.shape)
        if verbose:
            residuals, smoothness_penalty = obj.parts(result.x)
            say("\nresiduals", residuals)
            say("smoothness_penalty", smoothness_penalty)
        history.append(x[::-1, ...] if indexing == "yx" else x)

    # sector field -> pixel grid of the level frame (vet.py:621-630), then everything the reference
    # undoes on the host -- level padding, axis order, global padding -- as one view of the device
    # tensor; a single D2H of the result
    M, N = level.shape[1], level.shape[2]
    bi, bj = x.shape[1], x.shape[2]
    full = _zoom_on_device(x, int(round(bi * (M / bi))), int(round(bj * (N / bj))))
    (pi0, pi1), (pj0, pj1) = pads
    full = full[:, pi0: M - pi1, pj0: N - pj1]
    if indexing == "yx":
        full = full.flip(0)
    if padding > 0:
        full = full[:, padding:-padding, padding:-padding]
    field = _device.remember_result(_device.to_host(full.contiguous()), full.contiguous()) \
        if full.numel() else full.cpu().numpy()
    if intermediate_steps:
        return field, history
    return field
