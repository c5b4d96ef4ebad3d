"""B200 Variational Echo Tracking -- drop-in for ``pysteps.motion.vet``
(pysteps/motion/vet.py:93-648).

The reference splits VET into a Python driver (masking, padding, sector pyramid,
``scipy.optimize.minimize`` CG loop, ``scipy.ndimage.zoom`` upsampling) and a native
extension (``_vet.pyx``: ``_warp`` and ``_cost_function``).  Here the driver is mirrored in
Python, the optimiser stays SciPy's, and the native extension is ``csrc/vet.cu``: each
cost / gradient evaluation of the ~1000 the optimiser requests is one fused kernel over the
device-resident image pair plus a tiny finalising kernel (``b200_vet_cost``); per evaluation
the host sends <= 2*32*32 sector displacements and reads back a scalar or that many
gradient values.  The final full-resolution zoom runs on the device (``b200_zoom_bilinear``).
"""
import numpy
import torch
from numpy.ma.core import MaskedArray
from scipy.ndimage import zoom
from scipy.optimize import minimize

from .. import _device, _lib


def round_int(scalar):
    """Round number to nearest integer (vet.py:41-45)."""
    return int(numpy.round(scalar))


def ceil_int(scalar):
    """Round number up to the nearest integer (vet.py:48-52)."""
    return int(numpy.ceil(scalar))


def get_padding(dimension_size, sectors):
    """Padding before/after so that `sectors` divides the dimension (vet.py:55-90)."""
    reminder = dimension_size % sectors
    if reminder != 0:
        pad = sectors - reminder
        pad_before = pad // 2
        if pad % 2 == 0:
            pad_after = pad_before
        else:
            pad_after = pad_before + 1
        return pad_before, pad_after
    return 0, 0


def morph(image, displacement, gradient=False):
    """Morph an image by a displacement field (vet.py:93-153 -> _vet._warp).
    Returns (image, mask) or (image, mask, gradient) as NumPy arrays."""
    _device.require_cuda()
    if not isinstance(image, MaskedArray):
        _mask = numpy.zeros_like(image, dtype="int8")
    else:
        _mask = numpy.asarray(numpy.ma.getmaskarray(image), dtype="int8", order="C")
    _image = numpy.asarray(image, dtype="float64", order="C")
    _displacement = numpy.asarray(displacement, dtype="float64", order="C")
    nx, ny = _image.shape
    d_img = _device.to_device(_image)
    d_mask = _device.to_device(_mask)
    d_disp = _device.to_device(_displacement)
    out = torch.empty((nx, ny), dtype=torch.float64, device="cuda")
    omask = torch.empty((nx, ny), dtype=torch.int8, device="cuda")
    grad = torch.empty((2, nx, ny), dtype=torch.float64, device="cuda") if gradient else None
    _lib.call("b200_vet_warp", d_img.data_ptr(), d_mask.data_ptr(), d_disp.data_ptr(), nx, ny,
              out.data_ptr(), omask.data_ptr(), _device.ptr(grad), _device.stream_ptr())
    if gradient:
        return out.cpu().numpy(), omask.cpu().numpy(), grad.cpu().numpy()
    return out.cpu().numpy(), omask.cpu().numpy()


class _DeviceImages:
    """Image stack and mask of one minimisation level, resident on the device."""

    def __init__(self, input_images, mask):
        self.shape = input_images.shape
        self.images = _device.to_device(numpy.ascontiguousarray(input_images, dtype=numpy.float64))
        self.mask = _device.to_device(numpy.ascontiguousarray(mask, dtype=numpy.int8))
        self._sd = {}
        self._out = {}

    def evaluate(self, sector_displacement_2d, pair, smooth_gain, gradient):
        """_cost_function(sector_displacement, images[pair[0]], images[pair[1]], mask, ...)."""
        shp = tuple(sector_displacement_2d.shape)
        if shp not in self._sd:
            self._sd[shp] = torch.empty(shp, dtype=torch.float64, device="cuda")
            self._out[shp] = (torch.empty(2, dtype=torch.float64, device="cuda"),
                              torch.empty(shp, dtype=torch.float64, device="cuda"))
        sd = self._sd[shp]
        sd.copy_(torch.from_numpy(numpy.ascontiguousarray(sector_displacement_2d, dtype=numpy.float64)))
        out = self._out[shp][1 if gradient else 0]
        nx, ny = int(self.shape[1]), int(self.shape[2])
        _lib.call("b200_vet_cost", sd.data_ptr(), self.images[pair[0]].data_ptr(),
                  self.images[pair[1]].data_ptr(), self.mask.data_ptr(), int(shp[1]), int(shp[2]), nx, ny,
                  float(smooth_gain), 1 if gradient else 0, out.data_ptr(), _device.stream_ptr())
        return out.cpu().numpy()


def vet_cost_function_gradient(*args, **kwargs):
    """Gradient of the VET cost function (vet.py:156-162)."""
    kwargs["gradient"] = True
    return vet_cost_function(*args, **kwargs)


def vet_cost_function(sector_displacement_1d, input_images, blocks_shape, mask, smooth_gain,
                      debug=False, gradient=False):
    """VET cost function / gradient (vet.py:165-299).  `input_images` may be the NumPy stack
    of the reference or a device-resident `_DeviceImages` (what `vet` passes, so that the
    images are uploaded once per minimisation level and not once per evaluation)."""
    _device.require_cuda()
    if not isinstance(input_images, _DeviceImages):
        input_images = _DeviceImages(numpy.asarray(input_images), numpy.asarray(mask))
    sector_displacement_2d = numpy.asarray(sector_displacement_1d).reshape(*((2,) + tuple(blocks_shape)))
    if sector_displacement_2d.shape[1] < 2 or sector_displacement_2d.shape[2] < 2:
        raise NotImplementedError("pysteps_b200 VET: at least 2 x 2 sectors are required")
    if (input_images.shape[1] % sector_displacement_2d.shape[1] != 0
            or input_images.shape[2] % sector_displacement_2d.shape[2] != 0):
        raise ValueError("Error computing cost function.\n",
                         "The number of sectors don't divide the image size")
    if input_images.shape[0] == 3:
        three_times = True
        pairs = ((1, 2), (0, 1))  # (center, next), then (previous, center)
    else:
        three_times = False
        pairs = ((0, 1),)
    if gradient:
        gradient_values = input_images.evaluate(sector_displacement_2d, pairs[0], smooth_gain, True)
        if three_times:
            gradient_values = gradient_values + input_images.evaluate(
                sector_displacement_2d, pairs[1], smooth_gain, True)
        return gradient_values.ravel()
    residuals, smoothness_penalty = input_images.evaluate(sector_displacement_2d, pairs[0], smooth_gain, False)
    if three_times:
        _residuals, _smoothness = input_images.evaluate(sector_displacement_2d, pairs[1], smooth_gain, False)
        residuals += _residuals
        smoothness_penalty += _smoothness
    if debug:
        print("\nresiduals", residuals)
        print("smoothness_penalty", smoothness_penalty)
    return residuals + smoothness_penalty


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

    if verbose:
        def debug_print(*args, **kwargs):
            print(*args, **kwargs)
    else:
        def debug_print(*args, **kwargs):
            del args
            del kwargs

    if options is None:
        options = dict()
    else:
        options = dict(options)

    options.setdefault("eps", 0.1)
    options.setdefault("gtol", 0.1)
    options.setdefault("maxiter", 100)
    options.setdefault("disp", False)
    optimization_method = options.pop("method", "CG")

    pad_i = None
    pad_j = None
    sectors_in_i = None
    sectors_in_j = None

    debug_print("Running VET algorithm")

    valid_indexing = ["yx", "xy", "ij"]
    if indexing not in valid_indexing:
        raise ValueError(
            "Invalid indexing values: {0}\n".format(indexing)
            + "Supported values: {0}".format(str(valid_indexing))
        )

    if not isinstance(input_images, MaskedArray):
        input_images = numpy.ma.masked_invalid(input_images)
    else:
        input_images = input_images.copy()  # the reference writes into .data below

    mask = numpy.ma.getmaskarray(input_images)

    if padding > 0:
        padding_tuple = ((0, 0), (padding, padding), (padding, padding))
        input_images_data = numpy.pad(numpy.ma.getdata(input_images), padding_tuple, "constant",
                                      constant_values=numpy.nan)
        mask = numpy.pad(mask, padding_tuple, "constant", constant_values=True)
        input_images = numpy.ma.MaskedArray(data=input_images_data, mask=mask)

    input_images.data[mask] = 0  # Remove any Nan from the raw data

    mask = numpy.asarray(numpy.any(mask, axis=0), dtype="int8", order="C")
    input_images = numpy.asarray(input_images.data, dtype="float64", order="C")

    sectors = numpy.asarray(sectors, dtype="int", order="C")
    if sectors.ndim == 1:
        new_sectors = numpy.zeros((2,) + sectors.shape, dtype="int", order="C") + sectors.reshape(
            (1, sectors.shape[0]))
        sectors = new_sectors
    elif sectors.ndim > 2 or sectors.ndim < 1:
        raise ValueError(
            "Incorrect sectors dimensions.\n"
            + "Only 1D or 2D arrays are supported to define"
            + "the number of sectors used in"
            + "the scaling procedure"
        )

    sectors[0, :].sort()
    sectors[1, :].sort()

    first_guess_shape = (2, int(sectors[0, 0]), int(sectors[1, 0]))
    if first_guess is None:
        first_guess = numpy.zeros(first_guess_shape, order="C")
    else:
        if first_guess.shape != first_guess_shape:
            raise ValueError(
                "The shape of the initial guess do not match the number of "
                + "sectors of the first scaling guess\n"
                + "first_guess.shape={}\n".format(str(first_guess.shape))
                + "Expected shape={}".format(str(first_guess_shape))
            )
        else:
            first_guess = numpy.asarray(first_guess, order="C", dtype="float64")

    scaling_guesses = list()
    previous_sectors_in_i = sectors[0, 0]
    previous_sectors_in_j = sectors[1, 0]
    _shape = input_images.shape
    device_cache = {}

    for n, (sectors_in_i, sectors_in_j) in enumerate(zip(sectors[0, :], sectors[1, :])):
        pad_i = get_padding(input_images.shape[1], sectors_in_i)
        pad_j = get_padding(input_images.shape[2], sectors_in_j)

        if (pad_i != (0, 0)) or (pad_j != (0, 0)):
            _input_images = numpy.pad(input_images, ((0, 0), pad_i, pad_j), "edge")
            _mask = numpy.pad(mask, (pad_i, pad_j), "constant", constant_values=1)
            _mask = numpy.ascontiguousarray(_mask)
        else:
            _input_images = input_images
            _mask = mask
        _shape = _input_images.shape

        # one upload per distinct padding (the images of a level do not change during CG)
        key = (pad_i, pad_j)
        if key not in device_cache:
            device_cache[key] = _DeviceImages(_input_images, _mask)
        dev_images = device_cache[key]

        sector_shape = (_shape[1] // sectors_in_i, _shape[2] // sectors_in_j)
        debug_print("original image shape: " + str(input_images.shape))
        debug_print("padded image shape: " + str(_shape))
        debug_print("padded template_image image shape: " + str(_shape))
        debug_print("\nNumber of sectors: {0:d},{1:d}".format(sectors_in_i, sectors_in_j))
        debug_print("Sector Shape:", sector_shape)

        if n > 0:
            first_guess = zoom(
                first_guess,
                (1, sectors_in_i / previous_sectors_in_i, sectors_in_j / previous_sectors_in_j),
                order=1, mode="nearest",
            )

        debug_print("Minimizing")
        result = minimize(
            vet_cost_function,
            first_guess.flatten(),
            jac=vet_cost_function_gradient,
            args=(dev_images, (sectors_in_i, sectors_in_j), _mask, smooth_gain),
            method=optimization_method,
            options=options,
        )
        first_guess = result.x.reshape(*first_guess.shape)

        if verbose:
            vet_cost_function(result.x, dev_images, (sectors_in_i, sectors_in_j), _mask, smooth_gain,
                              debug=True)
        if indexing == "yx":
            scaling_guesses.append(first_guess[::-1, ...])
        else:
            scaling_guesses.append(first_guess)

        previous_sectors_in_i = sectors_in_i
        previous_sectors_in_j = sectors_in_j

    # final zoom to the image grid (vet.py:621-630) on the device
    ni, nj = int(_shape[1]), int(_shape[2])
    oh = int(round(first_guess.shape[1] * (ni / sectors_in_i)))
    ow = int(round(first_guess.shape[2] * (nj / sectors_in_j)))
    d_fg = _device.to_device(numpy.ascontiguousarray(first_guess, dtype=numpy.float64))
    d_full = torch.empty((2, oh, ow), dtype=torch.float64, device="cuda")
    _lib.call("b200_zoom_bilinear", d_fg.data_ptr(), 2, int(first_guess.shape[1]), int(first_guess.shape[2]),
              oh, ow, d_full.data_ptr(), _device.stream_ptr())
    first_guess = _device.to_host(d_full)

    first_guess = first_guess[:, pad_i[0]: ni - pad_i[1], pad_j[0]: nj - pad_j[1]]
    if indexing == "yx":
        first_guess = first_guess[::-1, ...]
    if padding > 0:
        first_guess = first_guess[:, padding:-padding, padding:-padding]
    if intermediate_steps:
        return first_guess, scaling_guesses
    return first_guess
