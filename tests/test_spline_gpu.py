"""GPU parity tests of interp_order 0 and 2..5 (csrc/spline.cu) against the reference-generated
goldens and the oracle (kernel bodies and host logic are additionally run on the CPU:
tests/test_kernel_bodies.py, tests/test_host_logic_sl.py)."""
import os

import numpy as np
import pytest
from conftest import assert_bits_equal

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def extrap():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a GPU"
    import pysteps_b200
    return pysteps_b200.extrapolation.get_method("semilagrangian")


def test_reference_goldens(extrap):
    from sl_cases import SPLINE_CASES, build_case
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "sl_golden.npz"))
    for name in SPLINE_CASES:
        args, kwargs = build_case(name)
        res = extrap(*args, **kwargs)
        out, disp = res if isinstance(res, tuple) else (res, None)
        assert_bits_equal(out, golden[name + "/out"], name + " output")
        if disp is not None:
            assert_bits_equal(disp, golden[name + "/disp"], name + " displacement")


@pytest.mark.parametrize("order", [0, 2, 3, 4, 5])
@pytest.mark.parametrize("mode", ["constant", "nearest"])
def test_randomised_vs_oracle(extrap, order, mode):
    from oracle import semilagrangian as ora
    rng = np.random.default_rng(40 + order)
    for case in range(25):
        m, n = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        dtype = rng.choice([np.float64, np.float32])
        P = (rng.standard_normal((m, n)) * 5).astype(dtype)
        kw = {"interp_order": order, "map_coordinates_mode": mode}
        if rng.random() < 0.5:
            P[rng.random((m, n)) < 0.2] = np.nan
            P[0, 0] = 1.0
            kw["allow_nonfinite_values"] = True
        V = (rng.standard_normal((2, m, n)) * rng.choice([0.5, 3.0, 30.0])).astype(rng.choice([np.float64, np.float32]))
        if rng.random() < 0.4:
            kw["n_iter"] = int(rng.integers(0, 4))
        if rng.random() < 0.4:
            kw["return_displacement"] = True
        if rng.random() < 0.3:
            kw["displacement_prev"] = rng.standard_normal((2, m, n))
        outval = rng.choice([np.nan, 0.0, -15.0]) if rng.random() < 0.8 else "min"
        ts = int(rng.integers(1, 5)) if rng.random() < 0.5 else [0.5, 1.25, 3.0]
        got = extrap(P, V, ts, outval, **kw)
        want = ora.extrapolate(P, V, ts, outval, **kw)
        for a, b in zip(got if isinstance(got, tuple) else (got,), want if isinstance(want, tuple) else (want,)):
            assert_bits_equal(a, b, f"case {case} {(m, n)} {kw}")


def test_full_size_band_and_device_io(extrap):
    import torch
    from pysteps_b200 import _synthetic as syn
    m = n = 2048
    P = syn.rain_field(m, n, 0).astype(np.float32)
    V = syn.velocity_field(m, n, 0)
    full = extrap(P, V, 3, interp_order=3, map_coordinates_mode="nearest")
    band = extrap(P, V, 3, interp_order=3, map_coordinates_mode="nearest", b200_rows=(700, 1300))
    assert_bits_equal(band, full[:, 700:1300], "band")
    dev = extrap(torch.from_numpy(P).cuda(), torch.from_numpy(V).cuda(), 3, interp_order=3,
                 map_coordinates_mode="nearest")
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), full, equal_nan=True)
    # zero motion reproduces the field wherever it is wet (order 3 interpolates its samples to 1 ulp)
    still = extrap(P.astype(np.float64), np.zeros((2, m, n)), 1, interp_order=3)[0]
    assert np.abs(still - P).max() < 1e-9
