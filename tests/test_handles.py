"""CPU tests of the lazy-handle dispatch of pysteps_b200.noise.motion (no GPU: the device field
is replaced by a stand-in that records what would be launched)."""
import numpy as np
import pytest

from pysteps_b200 import _device
from pysteps_b200.noise import interface as noise_interface
from pysteps_b200.noise import motion as bps


class _FakeField:
    def __init__(self, V):
        self.host = V
        self.fp = bps._fingerprint(V)
        self.tensor = None
        self.shape = V.shape
        self.calls = []

    def run(self, a, b, vsf, what, n_nonfinite=None):
        self.calls.append((a, b, vsf, what))
        return ("device", what)


def test_generate_bps_coefficients_follow_the_reference_expression():
    V = np.ones((2, 5, 7))
    f = _FakeField(V)
    pert = bps._Perturbator(vsf=30.0, p_par=(10.88, 0.23, -7.68), p_perp=(5.76, 0.31, -2.72),
                            eps_par=-0.2, eps_perp=1.6, _field=f)
    h = bps.generate_bps(pert, 15.0)
    # noise/motion.py:177-180
    g_par = 10.88 * pow(15.0, 0.23) + -7.68
    g_perp = 5.76 * pow(15.0, 0.31) + -2.72
    assert isinstance(h, bps.Perturbation)
    assert (h.a, h.b, h.vsf) == (g_par * -0.2, g_perp * 1.6, 30.0)
    assert h.shape == (2, 5, 7) and h.ndim == 3 and h.dtype == np.float64 and len(h) == 2


def test_adding_the_perturbators_own_field_stays_lazy():
    V = np.ones((2, 5, 7))
    f = _FakeField(V)
    h = bps.Perturbation(f, 0.5, -0.25, 60.0)
    for s in (V + h, h + V):          # nowcasts/utils.py:449 spells it `velocity + pert`
        assert isinstance(s, bps.PerturbedVelocity)
        assert s.shape == V.shape and s.ndim == 3
    assert f.calls == []              # nothing launched yet
    assert (V + h).device_interleaved() == ("device", bps._FIELD_INTERLEAVED)
    assert (V + h).device_planar() == ("device", bps._FIELD_PLANAR)
    assert h.device_planar() == ("device", bps._PERTURBATION)
    assert [c[3] for c in f.calls] == [bps._FIELD_INTERLEAVED, bps._FIELD_PLANAR, bps._PERTURBATION]
    assert all(c[:3] == (0.5, -0.25, 60.0) for c in f.calls)


def test_a_field_rewritten_in_place_is_not_taken_for_the_initial_one():
    """velocity + generate_bps(...) is fused only while `velocity` still holds what initialize_bps
    saw; after an in-place edit the sum is materialised from the CURRENT array, as the reference
    would compute it."""
    V = np.ones((2, 5, 7))
    f = _FakeField(V)
    h = bps.Perturbation(f, 0.5, -0.25, 60.0)
    assert isinstance(V + h, bps.PerturbedVelocity)
    V[0, 0, 0] = 7.0
    f.run = lambda *a, **k: (_ for _ in ()).throw(AssertionError("materialise through np.asarray instead"))
    h.__class__ = type("P", (bps.Perturbation,), {"__array__": lambda self, dtype=None, copy=None: np.zeros((2, 5, 7))})
    s = V + h
    assert isinstance(s, np.ndarray) and s[0, 0, 0] == 7.0


def test_ndarray_defers_to_the_handle():
    # __array_ufunc__ = None makes ndarray.__add__ return NotImplemented instead of broadcasting
    # over the handle element by element
    V = np.ones((2, 3, 3))
    h = bps.Perturbation(_FakeField(V), 1.0, 1.0, 1.0)
    assert bps.Perturbation.__array_ufunc__ is None
    assert isinstance(V + h, bps.PerturbedVelocity)
    with pytest.raises(TypeError):
        np.add(V, h)


def test_perturbator_is_a_dict_with_lazy_unit_vectors():
    p = bps._Perturbator(vsf=1.0)
    assert p["vsf"] == 1.0 and isinstance(p, dict)
    with pytest.raises(KeyError):
        p["nope"]
    assert "V_par" not in p           # produced on first access only


def test_registry_and_defaults():
    init, gen = noise_interface.get_method("BPS")
    assert init is bps.initialize_bps and gen is bps.generate_bps
    assert noise_interface.get_method("bps_b200") == (init, gen)
    with pytest.raises(TypeError):
        noise_interface.get_method(None)
    with pytest.raises(ValueError, match="Unknown method"):
        noise_interface.get_method("nested")
    assert bps.get_default_params_bps_par() == (10.88, 0.23, -7.68)
    assert bps.get_default_params_bps_perp() == (5.76, 0.31, -2.72)
    # argument checks come before any device work (noise/motion.py:102-116)
    with pytest.raises(ValueError, match="three-dimensional"):
        bps.initialize_bps(np.ones((4, 4)), 1, 1)
    with pytest.raises(ValueError, match="first dimension"):
        bps.initialize_bps(np.ones((3, 4, 4)), 1, 1)
    with pytest.raises(ValueError, match="p_par"):
        bps.initialize_bps(np.ones((2, 4, 4)), 1, 1, p_par=(1,))
    with pytest.raises(ValueError, match="p_perp"):
        bps.initialize_bps(np.ones((2, 4, 4)), 1, 1, p_perp=(1,))


def test_fingerprint_notices_in_place_changes_at_sampled_positions():
    V = np.zeros((2, 64, 64))
    a = bps._fingerprint(V)
    assert bps._fingerprint(V) == a
    V[0, 0, 0] = 1.0
    assert bps._fingerprint(V) != a
    assert bps._fingerprint(V.copy())[2] != bps._fingerprint(V)[2]   # another buffer


def test_device_field_is_array_like_without_touching_the_device():
    class T:
        shape = (2, 3, 4)
        ndim = 3
        dtype = "torch.float64"
    d = _device.DeviceField(T())
    assert d.shape == (2, 3, 4) and d.ndim == 3 and d.dtype == np.float64 and len(d) == 2
