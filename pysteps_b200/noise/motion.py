"""B200 BPS motion perturbator -- drop-in for ``pysteps.noise.motion.initialize_bps`` /
``generate_bps`` (pysteps/noise/motion.py:55-180), fused with the advection call that
consumes it (pysteps/nowcasts/utils.py:448-458).

The reference materialises, per ensemble member and time step, a (2,m,n) float64
perturbation on the host (`generate_bps`), adds it to the motion field (`velocity + ...`) and
hands the sum to the extrapolator: 64 MB of host traffic plus -- with a GPU extrapolator -- a
64 MB upload, per member-step at 2048^2.  Here the motion field is uploaded once and shared by
all members; `generate_bps` returns a `Perturbation` handle holding two scalars; `velocity +
handle` returns a `PerturbedVelocity` handle; and `pysteps_b200` `extrapolate` turns that into
one kernel (`b200_bps_perturb_velocity`) that writes V + p straight into the layout the
trajectory kernel reads.  The values at the grid nodes are those of the reference, bit for bit.
Anything else that touches a handle (``np.asarray``, arithmetic with another array) gets the
materialised field, computed on the device.

The random draws (`randstate.laplace`, :126-127) stay on the host RNG stream, so a seeded
ensemble perturbs its members exactly like the reference.
"""
import weakref

import numpy as np
import torch

from .. import _device, _lib

_FIELD_INTERLEAVED, _FIELD_PLANAR, _PERTURBATION, _UNIT = 0, 1, 2, 3


def get_default_params_bps_par():
    """motion.py:43-46"""
    return (10.88, 0.23, -7.68)


def get_default_params_bps_perp():
    """motion.py:49-52"""
    return (5.76, 0.31, -2.72)


# ---------------------------------------------------------------------------------------------
class _BaseField:
    """The motion field of an ensemble, resident in HBM (planar (2,m,n), float32 or float64)."""

    def __init__(self, V):
        self.host = None if isinstance(V, torch.Tensor) else V
        self.fp = None if isinstance(V, torch.Tensor) else _fingerprint(V)
        if isinstance(V, torch.Tensor):
            t = V if V.dtype in (torch.float32, torch.float64) else V.to(torch.float64)
            self.tensor = _device.to_device(t)
        else:
            a = np.asarray(V)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)  # scipy.linalg.norm promotes integer input
            kept = _device.recall_result(a)  # the field a pysteps_b200 motion method returned
            self.tensor = kept if kept is not None else _device.to_device(a)
        self.shape = tuple(self.tensor.shape)
        # scipy.linalg.norm(check_finite=True), motion.py:134
        st = torch.empty(4, dtype=torch.float64, device="cuda")
        _lib.call("b200_field_stats", self.tensor.data_ptr(), _device.dtype_code(self.tensor.dtype),
                  self.tensor.numel(), st.data_ptr(), _device.stream_ptr())
        self.finite = float(st[0].item()) == 0.0

    def run(self, a, b, vsf, what, n_nonfinite=None):
        _, m, n = self.shape
        shape = (m, n, 2) if what == _FIELD_INTERLEAVED else (2, m, n)
        out = torch.empty(shape, dtype=torch.float64, device="cuda")
        _lib.call("b200_bps_perturb_velocity", self.tensor.data_ptr(), _device.dtype_code(self.tensor.dtype),
                  m, n, float(a), float(b), float(vsf), what, out.data_ptr(), _device.ptr(n_nonfinite),
                  _device.stream_ptr())
        return out


_fields = {}  # id(V) -> (weakref to V, fingerprint, _BaseField)


def _fingerprint(V):
    """Cheap guard against an array mutated in place between two initialize_bps calls."""
    a = np.asarray(V)
    flat = a.reshape(-1)
    step = max(1, flat.size // 4096)
    return (a.shape, a.dtype.str, a.__array_interface__["data"][0], flat[::step].tobytes())


def _base_field(V):
    """One upload per motion field: the members of an ensemble all pass the same array
    (nowcasts/steps.py:915-926)."""
    key = id(V)
    hit = _fields.get(key)
    if isinstance(V, torch.Tensor):
        fp = (tuple(V.shape), V.dtype, V.data_ptr(), V._version)  # _version counts in-place writes
    else:
        fp = _fingerprint(V)
    if hit is not None and hit[0]() is V and hit[1] == fp:
        return hit[2]
    bf = _BaseField(V)
    try:
        # the entry (and with it the device copy) goes when the host array does
        ref = weakref.ref(V, lambda _r, key=key, d=_fields: d.pop(key, None))
        if len(_fields) > 16:
            _fields.clear()
        _fields[key] = (ref, fp, bf)
    except TypeError:
        pass
    return bf


class _Perturbator(dict):
    """The dict `initialize_bps` returns.  "V_par" / "V_perp" (motion.py:138-141) are produced
    on first access -- the GPU path never needs them on the host."""

    def __missing__(self, key):
        if key in ("V_par", "V_perp"):
            unit = _device.to_host(self["_field"].run(0.0, 0.0, 1.0, _UNIT))
            self["V_par"] = unit
            self["V_perp"] = np.stack([-unit[1, :, :], unit[0, :, :]])
            return dict.__getitem__(self, key)
        raise KeyError(key)


class _Handle:
    """Common part of the two lazy (2,m,n) float64 fields."""
    __array_ufunc__ = None  # ndarray + handle -> handle.__radd__
    ndim = 3
    dtype = np.dtype(np.float64)

    def __array__(self, dtype=None, copy=None):
        a = _device.to_host(self.device_planar())
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return np.asarray(self)[idx]

    def __len__(self):
        return 2

    # any other arithmetic sees the materialised array
    def __sub__(self, o): return np.asarray(self) - o
    def __rsub__(self, o): return o - np.asarray(self)
    def __mul__(self, o): return np.asarray(self) * o
    def __rmul__(self, o): return o * np.asarray(self)
    def __truediv__(self, o): return np.asarray(self) / o
    def __neg__(self): return -np.asarray(self)


class Perturbation(_Handle):
    """Value of generate_bps(perturbator, t): (a*V_par + b*V_perp)/vsf, not materialised."""

    def __init__(self, field, a, b, vsf):
        self.field, self.a, self.b, self.vsf = field, a, b, vsf
        self.shape = field.shape

    def device_planar(self):
        return self.field.run(self.a, self.b, self.vsf, _PERTURBATION)

    def _add(self, other):
        if isinstance(other, torch.Tensor) and other is self.field.tensor:
            return PerturbedVelocity(self)
        # the host array the perturbator was initialised with -- unless it was rewritten in place
        # since (then the reference would add the perturbation to the CURRENT field: materialise)
        if other is self.field.host and _fingerprint(other) == self.field.fp:
            return PerturbedVelocity(self)
        if isinstance(other, torch.Tensor):
            return other + self.device_planar()
        return np.asarray(other) + np.asarray(self)

    __add__ = _add
    __radd__ = _add


class PerturbedVelocity(_Handle):
    """Value of ``velocity + generate_bps(perturbator, t)`` for the velocity the perturbator was
    initialised with; `pysteps_b200` extrapolate consumes it without materialising it."""

    def __init__(self, pert):
        self.pert = pert
        self.shape = pert.shape

    def device_interleaved(self, n_nonfinite=None):
        """(m,n,2) float64 field for the trajectory kernel; `n_nonfinite` (a device float64)
        receives the count of non-finite elements."""
        p = self.pert
        return p.field.run(p.a, p.b, p.vsf, _FIELD_INTERLEAVED, n_nonfinite)

    def device_planar(self):
        p = self.pert
        return p.field.run(p.a, p.b, p.vsf, _FIELD_PLANAR)

    def __add__(self, o): return np.asarray(self) + o
    __radd__ = __add__


# ---------------------------------------------------------------------------------------------
def initialize_bps(V, pixelsperkm, timestep, p_par=None, p_perp=None, randstate=None, seed=None):
    """Same contract as the reference (motion.py:55-141).  V may also be a CUDA tensor."""
    if len(V.shape) != 3:
        raise ValueError("V is not a three-dimensional array")
    if V.shape[0] != 2:
        raise ValueError("the first dimension of V is not 2")

    if p_par is None:
        p_par = get_default_params_bps_par()
    if p_perp is None:
        p_perp = get_default_params_bps_perp()

    if len(p_par) != 3:
        raise ValueError("the length of p_par is not 3")
    if len(p_perp) != 3:
        raise ValueError("the length of p_perp is not 3")

    _device.require_cuda()
    perturbator = _Perturbator()
    if randstate is None:
        randstate = np.random

    if seed is not None:
        randstate.seed(seed)

    eps_par = randstate.laplace(scale=1.0 / np.sqrt(2))
    eps_perp = randstate.laplace(scale=1.0 / np.sqrt(2))

    # scale factor for converting the unit of the advection velocities into km/h
    vsf = 60.0 / (timestep * pixelsperkm)

    field = _base_field(V)
    if not field.finite:
        raise ValueError("array must not contain infs or NaNs")  # scipy.linalg.norm, :134

    perturbator["randstate"] = randstate
    perturbator["vsf"] = vsf
    perturbator["p_par"] = p_par
    perturbator["p_perp"] = p_perp
    perturbator["eps_par"] = eps_par
    perturbator["eps_perp"] = eps_perp
    perturbator["_field"] = field
    return perturbator


def generate_bps(perturbator, t):
    """Same contract as the reference (motion.py:144-180); the result is a `Perturbation`
    handle (see the module docstring) -- ``np.asarray(result)`` is the reference's array."""
    vsf = perturbator["vsf"]
    p_par = perturbator["p_par"]
    p_perp = perturbator["p_perp"]
    eps_par = perturbator["eps_par"]
    eps_perp = perturbator["eps_perp"]

    g_par = p_par[0] * pow(t, p_par[1]) + p_par[2]
    g_perp = p_perp[0] * pow(t, p_perp[1]) + p_perp[2]

    return Perturbation(perturbator["_field"], g_par * eps_par, g_perp * eps_perp, vsf)
