"""GPU parity at the BASELINE.json sizes, inputs built exactly as bench.py builds them: the CUDA
path against the oracle (pinned to the reference: tests/test_oracle_*.py) on 2048^2 frames
(config[1]) and one 4096^2 / 24-leadtime composite (config[4]).  Bars: sparse vectors and
extrapolated fields / displacements BIT-identical, dense motion field <= 1e-12 at EVERY pixel."""
import numpy as np
import pytest
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a GPU"
    import pysteps_b200
    return (pysteps_b200.motion.get_method("lk"), pysteps_b200.extrapolation.get_method("semilagrangian"))


@pytest.mark.parametrize("seed,nframes", [(0, 2), (1, 2), (2, 2), (0, 3), (1, 3), (2, 3)])
def test_config1_lk_field_vs_oracle_2048(api, seed, nframes):
    from oracle import lucaskanade as ora
    from pysteps_b200 import _synthetic as syn
    lk, _ = api
    fr = syn.rain_frames(2048, 2048, nframes, seed)
    xy, uv = lk(fr, dense=False)
    oxy, ouv = ora.dense_lucaskanade(fr, dense=False)
    assert len(oxy) > 500 * (nframes - 1)  # the regime small frames never reach
    assert np.array_equal(xy, oxy) and np.array_equal(uv, ouv), "sparse vectors"
    V = lk(fr)
    Vo = ora.dense_lucaskanade(fr)
    assert V.shape == Vo.shape == (2, 2048, 2048)
    assert np.abs(V - Vo).max() <= 1e-12, "dense field, every pixel"


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("kind", ["smooth", "rotation", "lk"])
def test_config1_extrapolation_vs_oracle_2048(api, seed, kind):
    """12 leadtimes at 2048^2 with the bench's float32 precipitation field: outputs and final
    displacement bit-identical, for a smooth field, a solid-body rotation and the dense LK field."""
    from oracle import semilagrangian as ora
    from pysteps_b200 import _synthetic as syn
    lk, extrap = api
    fr = syn.rain_frames(2048, 2048, 2, seed)
    P = fr[-1].astype(np.float32)
    if kind == "lk":
        V = lk(fr)
    else:
        V = syn.velocity_field(2048, 2048, seed, kind)
        if kind == "rotation":
            V = V * 2.0
    got, gd = extrap(P, V, 12, return_displacement=True)
    want, wd = ora.extrapolate(P, V, 12, return_displacement=True)
    assert_bits_equal(got, want, f"{kind} seed {seed}: 12 leadtimes")
    assert_bits_equal(gd, wd, f"{kind} seed {seed}: displacement")


def test_config4_composite_4096_vs_oracle(api):
    from oracle import lucaskanade as ora_lk
    from oracle import semilagrangian as ora_sl
    from pysteps_b200 import _synthetic as syn
    lk, extrap = api
    fr = syn.rain_frames(4096, 4096, 2, 0)
    V = lk(fr)
    Vo = ora_lk.dense_lucaskanade(fr)
    assert np.abs(V - Vo).max() <= 1e-12
    P = fr[-1].astype(np.float32)
    got = extrap(P, Vo, 24)
    want = ora_sl.extrapolate(P, Vo, 24)
    assert_bits_equal(got, want, "4096^2, 24 leadtimes")
    # row bands (the multi-GPU partitioning of this config) are the rows of the full result
    band = extrap(P, Vo, 24, b200_rows=(1024, 1536))
    assert_bits_equal(band, want[:, 1024:1536], "row band")


def test_config2_vet_2048_vs_oracle():
    """BASELINE config[2]: VET at 2048^2.  One cost / gradient evaluation at every level's sector
    grid against the oracle (relative 1e-12), and the optimised field after a bounded number of CG
    iterations within 1e-6 px of the oracle run with the same options."""
    import pysteps_b200
    from oracle import vet as ora
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.motion import vet as b200_vet
    fr = syn.rain_frames(2048, 2048, 2, 0)
    mask = np.zeros((2048, 2048), np.int8)
    rng = np.random.default_rng(3)
    for bs in ((2, 2), (4, 4), (16, 16), (32, 32)):
        x = rng.normal(size=(2,) + bs).ravel() * 2.0
        c = b200_vet.vet_cost_function(x, fr, bs, mask, 1e6)
        g = b200_vet.vet_cost_function_gradient(x, fr, bs, mask, 1e6)
        co = ora.vet_cost_function(x, fr, bs, mask, 1e6)
        go = ora.vet_cost_function_gradient(x, fr, bs, mask, 1e6)
        assert abs(c - co) <= 1e-12 * abs(co), bs
        assert np.abs(g - go).max() <= 1e-12 * np.abs(go).max(), bs
    opts = {"maxiter": 4}
    V = pysteps_b200.motion.get_method("vet")(fr, verbose=False, options=opts)
    Vo = ora.vet(fr, verbose=False, options=opts)
    assert V.shape == Vo.shape == (2, 2048, 2048)
    assert np.abs(V - Vo).max() <= 1e-6
