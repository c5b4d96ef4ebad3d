"""CPU test of the HOST logic of pysteps_b200.motion.proesmans (frame checks, min/max scaling, NaN
and float32 handling, return conventions) with the C ABI emulated by the kernels' own bodies
compiled for the host (tests/cpu_abi.py -> tests/host_kernels): results must equal the oracle bit
for bit and the reference's stored outputs to the tolerance of tests/test_oracle_proesmans.py."""
import os

import numpy as np
import pytest

import cpu_abi
from oracle import proesmans as ora
from proesmans_cases import CASES, build_case


@pytest.fixture
def enabled():
    return None


def test_frame_checks_come_first():
    from pysteps_b200.motion import get_method
    with cpu_abi.emulated():
        with pytest.raises(IndexError):
            get_method("proesmans")(np.zeros((1, 16, 16)))
        # the frame checks come first, as in the reference (decorators.check_input_frames)
        with pytest.raises(ValueError, match="dimension mismatch"):
            get_method("proesmans")(np.zeros((16, 16)))
        with pytest.raises(ValueError, match="frames 3 mismatch"):
            get_method("proesmans_b200")(np.zeros((3, 16, 16)))


@pytest.mark.parametrize("name", CASES)
def test_shim_matches_oracle_and_reference_golden(name, enabled):
    from pysteps_b200.motion.proesmans import proesmans
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "proesmans_golden.npz"))
    frames, kw = build_case(name)
    with cpu_abi.emulated():
        if frames.shape[0] != 2:
            with pytest.raises(ValueError) as e:
                proesmans(frames, **kw)
            assert str(e.value) == str(golden[name + "/error"])
            frames = frames[-2:]
        adv, q = proesmans(frames, full_output=True, **kw)
        field = proesmans(frames, **kw)
    want_adv, want_q = ora.proesmans(frames, full_output=True, **kw)
    assert isinstance(adv, np.ndarray) and adv.dtype == np.float64 and adv.shape == want_adv.shape
    assert np.array_equal(adv, want_adv) and np.array_equal(q, want_q) and np.array_equal(field, want_adv[0])
    assert np.abs(adv - golden[name + "/advfield"]).max() <= 1e-8
    assert np.abs(q - golden[name + "/quality"]).max() <= 1e-7


def test_dtypes_nan_and_unsupported_options(enabled):
    from pysteps_b200.motion.proesmans import proesmans
    from pysteps_b200 import _synthetic as syn
    fr = syn.rain_frames(48, 40, 2, 4, dx=1, dy=1)
    with cpu_abi.emulated():
        # integer frames are promoted like (im - im_min) / ... does in the reference
        ints = np.rint(fr).astype(np.int64)
        assert np.array_equal(proesmans(ints, num_iter=5, num_levels=3), ora.proesmans(ints, num_iter=5, num_levels=3))
        # float32 frames hit the float64 memoryview of the reference extension
        with pytest.raises(ValueError, match="Buffer dtype mismatch"):
            proesmans(fr.astype(np.float32))
        # NaN: np.min / np.max propagate it, so the frames are not rescaled
        nan = fr.copy()
        nan[0, 3, 3] = np.nan
        got = proesmans(nan, num_iter=3, num_levels=2)
        want = ora.proesmans(nan, num_iter=3, num_levels=2)
        assert np.array_equal(got, want, equal_nan=True)
        # Gaussian pre-filter (scipy.ndimage.gaussian_filter restated, bit for bit)
        for std in (0.7, 2.0):
            got = proesmans(fr, filter_std=std, num_iter=5, num_levels=3, full_output=True)
            want = ora.proesmans(fr, filter_std=std, num_iter=5, num_levels=3, full_output=True)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        with pytest.raises(NotImplementedError, match="filter_std"):
            proesmans(fr, filter_std=20.0)
        with pytest.raises(NotImplementedError, match="empty pyramid level"):
            proesmans(fr, num_levels=9)
