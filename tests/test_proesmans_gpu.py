"""GPU parity tests of the Proesmans method (csrc/proesmans.cu) against the reference-generated
goldens and the oracle (bodies / host logic also on the CPU: tests/test_kernel_bodies.py,
tests/test_host_logic_proesmans.py)."""
import os

import numpy as np
import pytest
from conftest import assert_bits_equal

from proesmans_cases import STRICT_CASES, build_case

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def proesmans():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a GPU"
    import pysteps_b200
    return pysteps_b200.motion.get_method("proesmans")


@pytest.mark.parametrize("name", STRICT_CASES)
def test_bit_identical_to_the_oracle(proesmans, name):
    """Same operations in the same (wavefront == raster) order, row-wise mean on both sides."""
    from oracle import proesmans as ora
    frames, kw = build_case(name)
    adv, q = proesmans(frames[-2:], full_output=True, **kw)
    want_adv, want_q = ora.proesmans(frames[-2:], full_output=True, **kw)
    assert_bits_equal(adv, want_adv, name + " advection fields")
    assert_bits_equal(q, want_q, name + " consistency maps")


def test_reference_goldens_within_tolerance(proesmans):
    from proesmans_cases import CASES
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "proesmans_golden.npz"))
    for name in CASES:
        frames, kw = build_case(name)
        adv, q = proesmans(frames[-2:], full_output=True, **kw)
        assert np.abs(adv - golden[name + "/advfield"]).max() <= 1e-8, name
        assert np.abs(q - golden[name + "/quality"]).max() <= 1e-7, name


def test_full_size_recovers_translation_and_device_io(proesmans):
    import torch
    from pysteps_b200 import _synthetic as syn
    fr = syn.rain_frames(1024, 1024, 2, 0, dx=3, dy=-2)
    V = proesmans(fr, num_iter=30)
    assert V.shape == (2, 1024, 1024) and np.isfinite(V).all()
    wet = fr[0] > 1.0
    assert abs(np.median(V[0][wet]) - 3.0) < 0.5 and abs(np.median(V[1][wet]) + 2.0) < 0.5
    Vd = proesmans(torch.from_numpy(fr).cuda(), num_iter=30)
    assert Vd.is_cuda and np.array_equal(Vd.cpu().numpy(), V)
