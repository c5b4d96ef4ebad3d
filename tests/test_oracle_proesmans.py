"""CPU tests: the Proesmans oracle (oracle/proesmans_oracle.c, oracle/proesmans.py) is pinned
  * BIT FOR BIT against outputs of the reference SOURCE (pysteps/motion/proesmans.py +
    _proesmans.pyx) with the extension compiled without -ffast-math
    (tests/golden/gen_proesmans_strict_golden.py) -- ill-conditioned cases included;
  * to a tolerance against outputs of the extension compiled with the reference's own flags
    (-O3 -ffast-math, setup.py:27-28; tests/golden/gen_proesmans_golden.py): that build does not
    define its rounding, and the solver turns a flipped branch into a local change of up to a
    pixel, so only well-conditioned cases are compared closely;
  * against both builds live when they are still present in this container."""
import os

import numpy as np
import pytest
from conftest import assert_bits_equal

from oracle import proesmans as ora
from proesmans_cases import CASES, STRICT_CASES, build_case

TOL = 1e-8  # px, -ffast-math build, well-conditioned cases; flows are O(1) px


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "proesmans_golden.npz"))


@pytest.fixture(scope="module")
def strict_golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "proesmans_strict_golden.npz"))


@pytest.fixture
def raster_mean():
    ora.raster_order_mean(True)
    yield
    ora.raster_order_mean(False)


@pytest.mark.parametrize("name", STRICT_CASES)
def test_oracle_is_bit_identical_to_the_ieee_build_of_the_reference(name, strict_golden, raster_mean):
    frames, kw = build_case(name)
    adv, q = ora.proesmans(frames[-2:], full_output=True, **kw)
    assert_bits_equal(adv, strict_golden[name + "/advfield"], name + " advection fields")
    assert_bits_equal(q, strict_golden[name + "/quality"], name + " consistency maps")


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_fastmath_build_of_the_reference(name, golden):
    frames, kw = build_case(name)
    if frames.shape[0] != 2:
        with pytest.raises(ValueError) as e:
            ora.proesmans(frames, **kw)
        assert str(e.value) == str(golden[name + "/error"])
        frames = frames[-2:]
    adv, q = ora.proesmans(frames, full_output=True, **kw)   # default: row-wise mean, as the CUDA path
    ref_adv, ref_q = golden[name + "/advfield"], golden[name + "/quality"]
    assert adv.shape == ref_adv.shape and q.shape == ref_q.shape
    assert np.abs(adv - ref_adv).max() <= TOL and np.abs(q - ref_q).max() <= 10 * TOL
    assert np.array_equal(ora.proesmans(frames, **kw), adv[0])
    if name != "constant_32x32":
        assert np.abs(ref_adv).max() > 0.3   # the cases do exercise the solver


def test_gaussian_prefilter_against_live_reference(raster_mean):
    ref = _live("/tmp/proes_strict")
    frames, _ = build_case("default_96x128")
    for std in (0.8, 2.5):
        a, qa = ref(frames, filter_std=std, num_iter=10, num_levels=3, full_output=True)
        b, qb = ora.proesmans(frames, filter_std=std, num_iter=10, num_levels=3, full_output=True)
        assert_bits_equal(b, a, f"filter_std {std}")
        assert_bits_equal(qb, qa, f"filter_std {std} quality")


def test_argument_errors():
    with pytest.raises(ValueError, match="dimension mismatch"):
        ora.proesmans(np.zeros((8, 8)))
    with pytest.raises(ValueError, match="frames 3 mismatch"):
        ora.proesmans(np.zeros((3, 8, 8)))


def _live(build):
    """The reference's proesmans() with the extension of `build`, the extension loaded straight
    from its file so that both builds can live in one process."""
    import glob
    import importlib.util
    from _refimport import available
    hits = glob.glob(os.path.join(build, "pysteps", "motion", "_proesmans*.so"))
    if not (available() and hits):
        pytest.skip(f"reference extension not built in {build}")
    try:
        spec = importlib.util.spec_from_file_location("pysteps.motion._proesmans", hits[0])
        ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ext)
    except ImportError:
        pytest.skip("reference extension not importable")

    def proesmans(input_images, lam=50.0, num_iter=100, num_levels=6, filter_std=0.0, full_output=False):
        # pysteps/motion/proesmans.py:73-94
        from scipy.ndimage import gaussian_filter
        im = np.stack([input_images[-2, :, :].copy(), input_images[-1, :, :].copy()])
        im_min, im_max = np.min(im), np.max(im)
        if im_max - im_min > 1e-8:
            im = (im - im_min) / (im_max - im_min) * 255.0
        if filter_std > 0.0:
            im[0, :, :] = gaussian_filter(im[0, :, :], filter_std)
            im[1, :, :] = gaussian_filter(im[1, :, :], filter_std)
        advfield, quality = ext._compute_advection_field(im, lam, num_iter, num_levels)
        return (advfield, quality) if full_output else advfield[0]

    return proesmans


def _random_cases(seed, count):
    from pysteps_b200 import _synthetic as syn
    rng = np.random.default_rng(seed)
    for _ in range(count):
        m, n = int(rng.integers(8, 160)), int(rng.integers(8, 160))
        lv = int(rng.integers(1, 6))
        if (m >> (lv - 1)) < 1 or (n >> (lv - 1)) < 1:
            continue   # an empty pyramid level: the reference reads out of bounds there
        fr = syn.rain_frames(m, n, 2, int(rng.integers(0, 100)), dx=int(rng.integers(-3, 4)), dy=int(rng.integers(-3, 4)))
        yield fr, dict(lam=float(rng.choice([5.0, 50.0, 1000.0])), num_iter=int(rng.choice([1, 7, 30, 100])),
                       num_levels=lv, full_output=True)


def test_live_ieee_build_bitwise(raster_mean):
    ref = _live("/tmp/proes_strict")
    for fr, kw in _random_cases(21, 25):
        a, qa = ref(fr, **kw)
        b, qb = ora.proesmans(fr, **kw)
        assert_bits_equal(b, a, f"{fr.shape} {kw}")
        assert_bits_equal(qb, qa, f"{fr.shape} {kw} quality")


def test_live_fastmath_build_statistically():
    ref = _live("/tmp/vetbuild")
    devs = []
    for fr, kw in _random_cases(22, 25):
        a, _ = ref(fr, **kw)
        b, _ = ora.proesmans(fr, **kw)
        devs.append(np.abs(a - b).max())
    assert np.median(devs) < 1e-9, sorted(devs)
