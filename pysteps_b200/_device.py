"""Device-memory plumbing.  PyTorch is used ONLY as a container for device /
pinned-host memory and as the owner of the CUDA stream; all arithmetic runs in
the kernels of ``libpysteps_b200.so``."""
import threading
import weakref

import numpy as np
import torch

from . import _lib

_init_lock = threading.Lock()
_initialised = False


def require_cuda():
    """Fail loudly when no GPU is usable -- there is no CPU path in this package."""
    global _initialised
    if _initialised:
        return
    with _init_lock:
        if _initialised:
            return
        if not torch.cuda.is_available():
            raise RuntimeError("pysteps_b200 needs a CUDA device (B200, sm_100a); none is "
                               "available and there is no CPU fallback.")
        torch.cuda.init()
        lib = _lib.load()
        torch.zeros(1, device="cuda")  # make sure the primary context exists
        sm = _lib.c_int(0)
        mj = _lib.c_int(0)
        mn = _lib.c_int(0)
        _lib.check(lib.b200_device_info(sm, mj, mn, None, 0))
        _initialised = True


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (what every C-ABI call is
    enqueued on).  The raw accessor is ~1 us; torch.cuda.current_stream() builds a Stream object
    (~15 us -- three per ensemble member-step add up)."""
    if _raw_stream is not None:
        try:
            return _raw_stream(torch.cuda.current_device())
        except Exception:  # noqa: BLE001 -- fall back to the public API
            pass
    return torch.cuda.current_stream().cuda_stream


def dtype_code(dt):
    if dt == torch.float32 or dt == np.float32:
        return _lib.F32
    if dt == torch.float64 or dt == np.float64:
        return _lib.F64
    raise TypeError(f"unsupported dtype {dt}")


def is_device_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


class DeviceField:
    """A result left in HBM on request (``b200_resident=True``): array-like enough to be stored
    and handed back (shape / ndim / dtype); ``np.asarray(x)`` downloads it once.  Passing it
    back to a pysteps_b200 function costs no transfer."""
    __slots__ = ("tensor", "_host")

    def __init__(self, tensor):
        self.tensor = tensor
        self._host = None

    shape = property(lambda self: tuple(self.tensor.shape))
    ndim = property(lambda self: self.tensor.ndim)
    dtype = property(lambda self: np.dtype(str(self.tensor.dtype).replace("torch.", "")))

    def __array__(self, dtype=None, copy=None):
        if self._host is None:
            self._host = to_host(self.tensor)
        return self._host if dtype is None else self._host.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return np.asarray(self)[idx]

    def __len__(self):
        return self.tensor.shape[0]


def to_device(a, dtype=None):
    """numpy array / torch tensor -> contiguous CUDA tensor (async H2D on the current stream;
    full PCIe speed when the host buffer is pinned)."""
    if isinstance(a, DeviceField):
        a = a.tensor
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.asarray(a)
        if not a.flags.c_contiguous or not a.flags.writeable:
            a = np.ascontiguousarray(a) if a.flags.writeable else np.array(a, order="C")
        t = torch.from_numpy(a)
    if not t.is_cuda:
        t = t.to("cuda", non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def to_host(t):
    """CUDA tensor -> numpy array backed by pinned memory (async D2H + one sync)."""
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def ptr(t):
    return None if t is None else t.data_ptr()


# ---- device copies of results that were handed out as NumPy arrays ---------------------------
# The plugin API is NumPy: motion.get_method("lk")(frames) returns an ndarray and the caller passes
# that same ndarray to extrapolation.get_method("semilagrangian") (nowcasts/steps.py:656-664 with
# the field of examples/plot_steps_nowcast.py:87).  The device tensor the array was downloaded from
# is remembered under the array's identity, so the second call does not upload 64 MB (2048^2) that
# are already in HBM.  Guards: the entry dies with the array (weakref), and a hit requires the same
# object, shape, dtype, buffer address AND an unchanged sample of its contents (every
# (size/16384)-th element) -- an array rewritten in place is uploaded again like any other array.
# A rewrite that touches none of the sampled elements is not seen: call forget_results() after
# editing a returned field in place at isolated pixels, or pass a copy.
_recent = {}
_recent_lock = threading.Lock()
_RECENT_MAX = 8


def _sample_fingerprint(a):
    flat = a.reshape(-1)
    step = max(1, flat.size // 16384)
    return (a.shape, a.dtype.str, a.__array_interface__["data"][0], flat[::step].tobytes())


def remember_result(host_array, tensor, finite=False):
    """Associate the NumPy result `host_array` with the device tensor it was copied from.
    finite=True: every element of `tensor` is finite by construction (a weighted mean of finite vectors);
    the private device copy is marked so that a consumer need not scan it again."""
    key = id(host_array)
    if finite:
        tensor._b200_finite = True  # the kept tensor is private to this module's table: nobody can edit it

    def _drop(_ref, key=key, d=_recent, lock=_recent_lock):  # globals are gone at interpreter exit
        with lock:
            d.pop(key, None)

    try:
        ref = weakref.ref(host_array, _drop)
    except TypeError:
        return host_array
    with _recent_lock:
        if len(_recent) >= _RECENT_MAX:
            _recent.pop(next(iter(_recent)))
        _recent[key] = (ref, _sample_fingerprint(host_array), tensor)
    return host_array


def recall_result(host_array):
    """The device tensor of a remembered, unmodified NumPy result, else None."""
    if not isinstance(host_array, np.ndarray):
        return None
    with _recent_lock:
        hit = _recent.get(id(host_array))
    if hit is None or hit[0]() is not host_array or not host_array.flags.c_contiguous:
        return None
    if hit[1] != _sample_fingerprint(host_array):
        with _recent_lock:
            _recent.pop(id(host_array), None)
        return None
    return hit[2]


def forget_results():
    """Drop every remembered device copy (after editing a returned array in place)."""
    with _recent_lock:
        _recent.clear()
