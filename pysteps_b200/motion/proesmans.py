"""B200 Proesmans et al. (1994) optical flow -- drop-in for ``pysteps.motion.proesmans.proesmans``
(pysteps/motion/proesmans.py:20-94) with the reference's native extension
(pysteps/motion/_proesmans.pyx) replaced by ``csrc/proesmans.cu``.

Parity: tests/test_proesmans_gpu.py (bit-identical to the IEEE build of the reference source on
the B200), tests/test_kernel_bodies.py and tests/test_host_logic_proesmans.py (kernel bodies and
host logic on the CPU).
"""
import numpy as np
import torch

from .. import _device, _lib


def proesmans(input_images, lam=50.0, num_iter=100, num_levels=6, filter_std=0.0, verbose=True,
              full_output=False):
    """Same contract as the reference (see its docstring, proesmans.py:29-71).  NumPy input ->
    NumPy results; CUDA tensor input -> results stay on the device."""
    # decorators.check_input_frames(2, 2): decorators.py:121-146
    if input_images.ndim != 3:
        raise ValueError(
            "input_images dimension mismatch.\n"
            f"input_images.shape: {str(tuple(input_images.shape))}\n"
            "(t, x, y ) dimensions expected"
        )
    if 2 < input_images.shape[0] > 2:
        raise ValueError(
            f"input_images frames {input_images.shape[0]} mismatch.\n"
            "Minimum frames: 2\n"
            "Maximum frames: 2\n"
        )
    if input_images.shape[0] < 2:
        # the reference's own check lets a single frame through and fails on im2 = input_images[-1]
        # of the missing pair with an IndexError (proesmans.py:73-77); say so before any upload
        raise IndexError("proesmans needs two input frames")
    del verbose  # Not used

    _device.require_cuda()
    on_device = _device.is_device_tensor(input_images)
    m, n = int(input_images.shape[1]), int(input_images.shape[2])
    if m >> (int(num_levels) - 1) < 1 or n >> (int(num_levels) - 1) < 1:
        raise NotImplementedError("pysteps_b200 proesmans: num_levels leaves an empty pyramid level")

    frames = input_images[-2:]
    if isinstance(frames, torch.Tensor):
        d_in = _device.to_device(frames if frames.dtype in (torch.float32, torch.float64) else frames.to(torch.float64))
    else:
        a = np.asarray(frames)
        d_in = _device.to_device(a if a.dtype in (np.float32, np.float64) else a.astype(np.float64))

    # proesmans.py:79-83 min / max scaling to 0..255 (np.min / np.max propagate NaN: no scaling then)
    st = torch.empty(4, dtype=torch.float64, device="cuda")
    _lib.call("b200_field_stats", d_in.data_ptr(), _device.dtype_code(d_in.dtype), d_in.numel(), st.data_ptr(),
              _device.stream_ptr())
    n_nonfinite, lo, hi, n_nan = st.cpu().tolist()
    if n_nan > 0:
        lo = hi = float("nan")
    if d_in.dtype == torch.float32:
        # the reference hands the (still float32) stack to a float64 Cython memoryview
        # (_proesmans.pyx:19), which rejects it
        raise ValueError("Buffer dtype mismatch, expected 'float64' but got 'float'")
    do_scale = bool(hi - lo > 1e-8)
    d_im = torch.empty((2, m, n), dtype=torch.float64, device="cuda")
    _lib.call("b200_proesmans_scale", d_in.data_ptr(), _device.dtype_code(d_in.dtype), 2 * m * n, float(lo),
              float(hi), int(do_scale), d_im.data_ptr(), _device.stream_ptr())

    if filter_std > 0.0:  # proesmans.py:85-87 scipy.ndimage.gaussian_filter on both frames
        sd = float(filter_std)
        radius = int(4.0 * sd + 0.5)
        if radius > 64:
            raise NotImplementedError("pysteps_b200 proesmans: filter_std > 16 is not implemented")
        # scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius)[::-1]
        x = np.arange(-radius, radius + 1)
        phi_x = np.exp(-0.5 / (filter_std * filter_std) * x ** 2)
        weights = np.ascontiguousarray((phi_x / phi_x.sum())[::-1], dtype=np.float64)
        d_f = torch.empty((2, m, n), dtype=torch.float64, device="cuda")
        for t in range(2):
            _lib.call("b200_gaussian_filter", d_im[t].data_ptr(), m, n, weights.ctypes.data_as(_lib.c_dp), radius,
                      d_f[t].data_ptr(), _device.stream_ptr())
        d_im = d_f

    d_adv = torch.empty((2, 2, m, n), dtype=torch.float64, device="cuda")
    d_q = torch.empty((2, m, n), dtype=torch.float64, device="cuda")
    _lib.call("b200_proesmans_field", d_im.data_ptr(), m, n, float(lam), int(num_iter), int(num_levels),
              d_adv.data_ptr(), d_q.data_ptr(), _device.stream_ptr())

    if not full_output:
        out = d_adv[0]
        return out if on_device else _device.to_host(out.contiguous())
    if on_device:
        return d_adv, d_q
    return _device.to_host(d_adv), _device.to_host(d_q)
