"""CPU test of the drop-in boundary against the real pysteps registries (only where
/root/reference exists; the GPU box has no pysteps)."""
import pytest


def test_register_into_reference_registries():
    import _refimport
    if not _refimport.available():
        pytest.skip("/root/reference not present")
    _refimport.import_reference()
    import importlib
    import sys
    from unittest.mock import MagicMock
    # the reference's Cython extensions are not built here; only the registries are needed
    for ext in ("pysteps.motion._proesmans", "pysteps.motion._vet"):
        sys.modules.setdefault(ext, MagicMock())
    ei = importlib.import_module("pysteps.extrapolation.interface")
    mi = importlib.import_module("pysteps.motion.interface")
    ni = importlib.import_module("pysteps.noise.interface")
    import pysteps_b200
    stock_sl = ei.get_method("semilagrangian")
    stock_lk = mi.get_method("lk")
    stock_bps = ni.get_method("bps")
    done = pysteps_b200.register()
    assert "extrapolation:semilagrangian_b200" in done and "motion:lk_b200" in done
    assert ei.get_method("semilagrangian_b200") is pysteps_b200.extrapolation.semilagrangian.extrapolate
    assert ei.get_method("SEMILAGRANGIAN_B200") is pysteps_b200.extrapolation.semilagrangian.extrapolate
    assert mi.get_method("lk_b200") is pysteps_b200.motion.lucaskanade.dense_lucaskanade
    assert mi.get_method("vet_b200") is pysteps_b200.motion.vet.vet
    assert ni.get_method("bps_b200") == (pysteps_b200.noise.motion.initialize_bps,
                                         pysteps_b200.noise.motion.generate_bps)
    assert ni.get_method("bps") is stock_bps
    # default registration leaves the stock names alone (pysteps/tests/test_interfaces.py:70,225-228)
    assert ei.get_method("semilagrangian") is stock_sl and mi.get_method("lk") is stock_lk
    try:
        pysteps_b200.register(override=True)
        assert ei.get_method("semilagrangian") is pysteps_b200.extrapolation.semilagrangian.extrapolate
        assert mi.get_method("LK") is pysteps_b200.motion.lucaskanade.dense_lucaskanade
        assert mi.get_method("vet") is pysteps_b200.motion.vet.vet
        assert ni.get_method("bps")[1] is pysteps_b200.noise.motion.generate_bps
    finally:
        ni._noise_methods["bps"] = stock_bps
        ei._extrapolation_methods["semilagrangian"] = stock_sl
        mi._methods["lk"] = stock_lk
        mi._methods["lucaskanade"] = stock_lk
        mi._methods["vet"] = importlib.import_module("pysteps.motion.vet").vet
