"""CPU tests: the VET oracle (oracle/vet.py + vet_oracle.c) against the committed outputs
of the reference vet.py driving the reference's own compiled _vet.pyx
(tests/golden/vet_golden.npz).  The reference extension is built with -ffast-math, so the
bar is relative 1e-9 on single evaluations and 1e-6 px on optimised fields."""
import os

import numpy as np
import pytest

from oracle import vet as ora
from vet_cases import EVAL_CASES, FIELD_CASES, eval_case, field_case


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "vet_golden.npz"))


@pytest.mark.parametrize("name", EVAL_CASES)
def test_cost_and_gradient(name, golden):
    sd, images, mask, gain = eval_case(name)
    shape = sd.shape[1:]
    c = ora.vet_cost_function(sd.ravel(), images, shape, mask, gain)
    g = ora.vet_cost_function_gradient(sd.ravel(), images, shape, mask, gain)
    assert abs(c - golden[name + "/cost"]) <= 1e-9 * abs(golden[name + "/cost"])
    ref = golden[name + "/grad"]
    assert np.abs(g - ref).max() <= 1e-9 * np.abs(ref).max()


@pytest.mark.parametrize("name", FIELD_CASES)
def test_optimised_field(name, golden):
    images, kw = field_case(name)
    field, steps = ora.vet(images, verbose=False, intermediate_steps=True, **kw)
    assert field.shape == golden[name + "/field"].shape
    assert np.abs(field - golden[name + "/field"]).max() < 1e-6
    for k, s in enumerate(steps):
        assert np.abs(s - golden[name + f"/step{k}"]).max() < 1e-6


def test_morph(golden):
    img = eval_case("eval_128x160_s4x4")[1][0]
    w, wm, wg = ora.warp(img, np.zeros(img.shape, np.int8), golden["morph/disp"], gradient=True)
    assert np.abs(w - golden["morph/image"]).max() < 1e-12
    assert np.array_equal(wm, golden["morph/mask"])
    assert np.abs(wg - golden["morph/grad"]).max() < 1e-12


def test_zoom_matches_scipy_bitwise():
    ndimage = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    for (c, h, w, oh, ow) in [(2, 2, 2, 4, 4), (2, 4, 4, 16, 16), (2, 16, 16, 32, 32), (2, 3, 5, 7, 64),
                              (2, 32, 16, 504, 1016)]:
        a = rng.normal(size=(c, h, w))
        z = ndimage.zoom(a, (1, oh / h, ow / w), order=1, mode="nearest")
        assert np.array_equal(z, ora.zoom_o1(a, oh, ow)), (c, h, w, oh, ow)


def test_sector_error():
    sd, images, mask, gain = eval_case("eval_96x96_s2x2")
    with pytest.raises(ValueError):
        ora.cost_function(np.zeros((2, 5, 5)), images[0], images[1], mask, gain)
