"""GPU parity tests of the fused BPS motion perturbation (pysteps_b200.noise.motion +
b200_bps_perturb_velocity) against the reference's stored outputs (tests/golden/bps_golden.npz),
its known-answer tests (pysteps/tests/test_noise_motion.py) and the oracle restatement.
Bit-exact: the perturbed field at the grid nodes is float64 arithmetic restated operation by
operation."""
import os

import numpy as np
import pytest
from conftest import assert_bits_equal
from numpy.testing import assert_array_almost_equal

from bps_cases import KINDS, LEADS, MEMBERS, TIMESTEP, fields

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a GPU"
    import pysteps_b200
    return pysteps_b200


@pytest.fixture(scope="module")
def golden_bps():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "bps_golden.npz"))


def test_known_answers(b200):
    # pysteps/tests/test_noise_motion.py:27-66, through the registry like nowcasts/steps.py:910
    init, gen = b200.noise.get_method("bps")
    v = np.ones((8, 8))
    velocity = np.stack([v, v])
    pert = init(velocity, 1, 1, seed=42)
    vv = np.ones((8, 8)) * np.sqrt(2) * 0.5
    for variable, expected in (("vsf", 60), ("eps_par", -0.2042896366299448), ("eps_perp", 1.6383482042624593),
                               ("p_par", (10.88, 0.23, -7.68)), ("p_perp", (5.76, 0.31, -2.72)),
                               ("V_par", np.stack([vv, vv])), ("V_perp", np.stack([-vv, vv]))):
        assert_array_almost_equal(pert[variable], expected)
    new_vv = gen(pert, 1)
    assert_array_almost_equal(new_vv, np.stack([v * -0.066401, v * 0.050992]))
    assert_array_almost_equal(b200.noise.motion.get_default_params_bps_par(), (10.88, 0.23, -7.68))
    assert_array_almost_equal(b200.noise.motion.get_default_params_bps_perp(), (5.76, 0.31, -2.72))


@pytest.mark.parametrize("kind", KINDS)
def test_member_loop_matches_reference_golden(b200, golden_bps, kind):
    """The loop body of pysteps/nowcasts/utils.py:440-458 with the B200 perturbator and
    extrapolator: handles end to end, displacement resident in HBM."""
    init, gen = b200.noise.get_method("bps_b200")
    extrap = b200.extrapolation.get_method("semilagrangian")
    P, V = fields(kind)
    for seed, kmpp in MEMBERS:
        pert = init(V, 1.0 / kmpp, TIMESTEP, randstate=np.random.RandomState(seed))
        key = f"{kind}/{seed}"
        assert_bits_equal(np.array([pert["eps_par"], pert["eps_perp"], pert["vsf"]]), golden_bps[key + "/eps"], "eps")
        assert_bits_equal(pert["V_par"], golden_bps[key + "/V_par"], "V_par")
        for t in LEADS:
            assert_bits_equal(np.asarray(gen(pert, t)), golden_bps[key + f"/pert_{t}"], f"pert {t}")
        disp = None
        for step in range(1, 4):
            Vp = V + gen(pert, step * TIMESTEP)
            assert isinstance(Vp, b200.noise.motion.PerturbedVelocity) and Vp.shape == V.shape and Vp.ndim == 3
            res, disp = extrap(P, Vp, [1.0], displacement_prev=disp, return_displacement=True, b200_resident=True)
            assert isinstance(res, np.ndarray) and not isinstance(disp, np.ndarray)
        assert_bits_equal(res[0], golden_bps[key + "/advected"], "advected")
        assert_bits_equal(np.asarray(disp), golden_bps[key + "/disp"], "disp")
        # the same loop with everything materialised on the host gives the same bits
        disp2 = None
        for step in range(1, 4):
            Vp = np.asarray(V + gen(pert, step * TIMESTEP))
            res2, disp2 = extrap(P, Vp, [1.0], displacement_prev=disp2, return_displacement=True)
        assert isinstance(disp2, np.ndarray)
        assert_bits_equal(res2[0], golden_bps[key + "/advected"], "advected (materialised)")
        assert_bits_equal(disp2, golden_bps[key + "/disp"], "disp (materialised)")


def test_handles_and_fallbacks(b200):
    import torch
    from oracle import noise_motion as ora
    init, gen = b200.noise.get_method("bps")
    rng = np.random.default_rng(8)
    for dt in (np.float64, np.float32, np.int64):
        V = (10 * rng.standard_normal((2, 37, 53))).astype(dt)
        V[:, 3:6, 4:9] = 0
        po = ora.initialize_bps(V, 0.5, 10.0, p_perp=(2.0, 0.5, 0.1), randstate=np.random.RandomState(5))
        pg = init(V, 0.5, 10.0, p_perp=(2.0, 0.5, 0.1), randstate=np.random.RandomState(5))
        assert po["eps_par"] == pg["eps_par"] and po["eps_perp"] == pg["eps_perp"] and po["vsf"] == pg["vsf"]
        assert_bits_equal(pg["V_perp"], po["V_perp"], "V_perp")
        for t in (0.0, 10.0, 55.0):
            h = gen(pg, t)
            assert h.shape == V.shape and h.dtype == np.float64
            assert_bits_equal(np.asarray(h), ora.generate_bps(po, t), "pert")
            assert_bits_equal(np.asarray(V + h), ora.perturbed_velocity(V, po, t), "V + pert")
            assert_bits_equal(np.asarray(h + V), ora.perturbed_velocity(V, po, t), "pert + V")
            other = rng.standard_normal(V.shape)   # not the perturbator's field -> plain array math
            got = other + h
            assert isinstance(got, np.ndarray)
            assert_bits_equal(got, other + ora.generate_bps(po, t), "other + pert")
            assert_bits_equal(2.0 * h, 2.0 * ora.generate_bps(po, t), "scalar * pert")
    # one device copy of the motion field per ensemble; an in-place change is noticed
    V = rng.standard_normal((2, 16, 16))
    p1, p2 = init(V, 1, 5), init(V, 1, 5)
    assert p1["_field"] is p2["_field"]
    V[0, 0, 0] += 1.0
    assert init(V, 1, 5)["_field"] is not p1["_field"]
    # CUDA tensor input stays on the device
    Vd = torch.from_numpy(V).cuda()
    pd = init(Vd, 1, 5, randstate=np.random.RandomState(1))
    ph = init(V, 1, 5, randstate=np.random.RandomState(1))
    assert_bits_equal(np.asarray(gen(pd, 5.0)), np.asarray(gen(ph, 5.0)), "tensor input")
    out = b200.extrapolation.get_method("semilagrangian")(torch.from_numpy(np.abs(V[0])).cuda(), Vd + gen(pd, 5.0), 2)
    assert torch.is_tensor(out) and out.is_cuda


def test_argument_errors(b200):
    init, _ = b200.noise.get_method("bps")
    v = np.ones((2, 4, 4))
    with pytest.raises(ValueError, match="three-dimensional"):
        init(v[0], 1, 1)
    with pytest.raises(ValueError, match="first dimension"):
        init(np.ones((3, 4, 4)), 1, 1)
    with pytest.raises(ValueError, match="p_par"):
        init(v, 1, 1, p_par=(1, 2))
    with pytest.raises(ValueError, match="p_perp"):
        init(v, 1, 1, p_perp=(1, 2, 3, 4))
    bad = v.copy()
    bad[0, 1, 1] = np.nan
    with pytest.raises(ValueError, match="infs or NaNs"):
        init(bad, 1, 1)
    # an overflowing perturbation makes the advection field non-finite: the extrapolator's own
    # check (semilagrangian.py:118-119) fires on the field the kernel produced
    _, gen = b200.noise.get_method("bps")
    pert = init(v, 1, 1, p_par=(1e308, 1.0, 0.0), seed=1)
    extrap = b200.extrapolation.get_method("semilagrangian")
    with pytest.raises(ValueError, match="velocity contains non-finite values"):
        extrap(np.ones((4, 4)), v + gen(pert, 100.0), 1)
    with pytest.raises(ValueError, match="velocity contains only non-finite values"):  # :122-123
        extrap(np.ones((4, 4)), v + gen(pert, 100.0), 1, allow_nonfinite_values=True)
    with pytest.raises(TypeError):
        b200.noise.get_method(None)
    with pytest.raises(ValueError, match="Unknown method"):
        b200.noise.get_method("parametric")


def test_full_size_properties(b200):
    """2048^2: a perturbation with zero amplitude is the unperturbed field bit for bit, and the
    fused path equals the materialised one."""
    from pysteps_b200 import _synthetic as syn
    init, gen = b200.noise.get_method("bps")
    extrap = b200.extrapolation.get_method("semilagrangian")
    m = n = 2048
    P = syn.rain_field(m, n, 0).astype(np.float32)
    V = syn.velocity_field(m, n, 0)
    pert = init(V, 1.0, 5.0, randstate=np.random.RandomState(3))
    zero = dict(pert)
    zero["eps_par"] = zero["eps_perp"] = 0.0
    zp = b200.noise.motion._Perturbator(zero)
    base = extrap(P, V, [1.0])
    assert_bits_equal(extrap(P, V + gen(zp, 5.0), [1.0]), base, "zero perturbation")
    fused = extrap(P, V + gen(pert, 30.0), [1.0])
    mat = extrap(P, np.asarray(V + gen(pert, 30.0)), [1.0])
    assert_bits_equal(fused, mat, "fused vs materialised")
    assert not np.array_equal(fused, base, equal_nan=True)


def test_batched_member_step_equals_member_by_member_calls(env_bps=None):
    """b200_sl_step_batched through extrapolate_members: every member's field and displacement
    bit-identical to its own single-step extrapolate call (float32 / float64 fields, 11 members =
    two batches of the kernel, three lead times with carried displacements, NaN inflow)."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a GPU"
    import pysteps_b200
    from pysteps_b200 import _synthetic as syn
    from pysteps_b200.extrapolation.semilagrangian import extrapolate, extrapolate_members
    init, gen = pysteps_b200.noise.get_method("bps")
    for dtype, vkind in ((np.float32, "smooth"), (np.float64, "rotation")):
        m, n, M = 200, 264, 11
        V = syn.velocity_field(m, n, 3, vkind) * (4.0 if vkind == "rotation" else 1.0)
        P = np.stack([syn.rain_field(m, n, 3 + j) for j in range(M)]).astype(dtype)
        perts = [init(V, 1.0, 5.0, randstate=np.random.RandomState(50 + j)) for j in range(M)]
        disp_b, disp_s = None, [None] * M
        for t in range(3):
            Vm = [V + gen(perts[j], (t + 1) * 5.0) for j in range(M)]
            out_b, disp_b = extrapolate_members(P, Vm, displacement_prev=disp_b)
            for j in range(M):
                o, disp_s[j] = extrapolate(P[j], Vm[j], [1.0], displacement_prev=disp_s[j], return_displacement=True)
                assert out_b.dtype == o.dtype
                assert_bits_equal(out_b[j], o[0], f"{dtype.__name__} t={t} member {j}")
                assert np.array_equal(np.asarray(disp_b)[j], np.asarray(disp_s[j])), (t, j)
        # device tensors in -> device tensors out
        dP = torch.from_numpy(P).cuda()
        o_d, d_d = extrapolate_members(dP, Vm, displacement_prev=None)
        assert o_d.is_cuda and d_d.is_cuda and tuple(d_d.shape) == (M, 2, m, n)
