"""The REFERENCE ITSELF against the CUDA path, on the B200: pysteps v1.21.3 travels to the GPU box
in compiled form (oracle/_ref, built by oracle/build_ref.py from /root/reference) and runs there
unmodified -- stand-alone functions, the ensemble member loop and a whole nowcasts.steps forecast,
once with its stock methods and once with the B200 methods registered over the stock names
(pysteps_b200.register(override=True)), real kernels both times."""
import contextlib
import io
import warnings

import numpy as np
import pytest
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test needs a GPU"
    from oracle import refimport
    if not refimport.available(extensions=True):
        pytest.skip("oracle/_ref has not been built (python oracle/build_ref.py where /root/reference exists)")
    return refimport


def _quiet(fn, *a, **k):
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        return fn(*a, **k)


@pytest.mark.parametrize("shape,nframes,seed", [((512, 512), 2, 0), ((512, 512), 3, 1), ((301, 417), 3, 2),
                                                ((1024, 1024), 2, 3)])
def test_dense_lucaskanade_equals_the_reference(ref, shape, nframes, seed):
    import pysteps_b200
    from pysteps_b200 import _synthetic as syn
    rlk = ref.ref_module("pysteps.motion.lucaskanade").dense_lucaskanade
    lk = pysteps_b200.motion.get_method("lk")
    fr = syn.rain_frames(shape[0], shape[1], nframes, seed)
    if seed == 2:
        fr = np.stack([syn.nan_disc(f, 0.1) for f in fr])
    rxy, ruv = _quiet(rlk, fr.copy(), dense=False)
    xy, uv = lk(fr.copy(), dense=False)
    assert np.array_equal(xy, rxy) and np.array_equal(uv, ruv), "sparse vectors"
    V, Vr = lk(fr.copy()), _quiet(rlk, fr.copy())
    assert V.shape == Vr.shape and V.dtype == Vr.dtype
    assert np.abs(V - Vr).max() <= 1e-12, "dense field at every pixel"


@pytest.mark.parametrize("case", range(6))
def test_extrapolate_equals_the_reference(ref, case):
    import pysteps_b200
    from pysteps_b200 import _synthetic as syn
    rsl = ref.ref_module("pysteps.extrapolation.semilagrangian").extrapolate
    sl = pysteps_b200.extrapolation.get_method("semilagrangian")
    m, n = ((256, 320), (511, 300), (128, 128), (400, 400), (200, 256), (333, 222))[case]
    P = syn.rain_field(m, n, case)
    V = syn.velocity_field(m, n, case, ("smooth", "rotation")[case % 2]) * (1.0 + case)
    kw = [dict(), dict(n_iter=3), dict(outval="min"), dict(interp_order=3, map_coordinates_mode="nearest"),
          dict(allow_nonfinite_values=True), dict(interp_order=0)][case]
    if case == 4:
        P = syn.nan_disc(P)
    ts = [4, 3, [0.5, 1.0, 2.5], 2, 3, 3][case]
    want, wd = rsl(P, V, ts, return_displacement=True, **kw)
    got, gd = sl(P, V, ts, return_displacement=True, **kw)
    assert_bits_equal(got, want, f"case {case} fields")
    assert_bits_equal(gd, wd, f"case {case} displacement")


def test_vet_close_to_the_reference_build(ref):
    """the reference extension is built with -ffast-math (setup.py:27-28): 1e-6 px"""
    import pysteps_b200
    from pysteps_b200 import _synthetic as syn
    rvet = ref.ref_module("pysteps.motion.vet").vet
    vet = pysteps_b200.motion.get_method("vet")
    fr = syn.rain_frames(256, 256, 2, 3)
    want = rvet(fr, verbose=False)
    got = vet(fr, verbose=False)
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-6


def test_steps_forecast_with_registered_b200_methods(ref):
    """Unmodified pysteps.nowcasts.steps.forecast (cascade, AR model, noise, BPS velocity
    perturbations, AR pre-alignment through the extrapolator, member loop), 3 members on 200^2,
    seeded: stock registries, then the same call with the B200 kernels registered OVER the stock
    names -- identical output."""
    import pysteps_b200
    from pysteps_b200 import _synthetic as syn
    steps = ref.ref_module("pysteps.nowcasts.steps")
    ex_if = ref.ref_module("pysteps.extrapolation.interface")
    mo_if = ref.ref_module("pysteps.motion.interface")
    no_if = ref.ref_module("pysteps.noise.interface")
    m = n = 200
    fr = syn.rain_frames(m, n, 3, 4, dx=2, dy=-1)
    R = np.where(fr > 0.1, 10 * np.log10(np.maximum(fr, 0.1)), -15.0)
    kw = dict(timesteps=4, n_ens_members=3, n_cascade_levels=4, precip_thr=-10.0, kmperpixel=1.0, timestep=5.0,
              noise_method="nonparametric", seed=42, num_workers=1)
    saved = (dict(ex_if._extrapolation_methods), dict(mo_if._methods), dict(no_if._noise_methods))
    try:
        V_stock = _quiet(mo_if.get_method("lk"), fr)
        want = _quiet(steps.forecast, R, V_stock, **kw)
        pysteps_b200.register(override=True)
        assert ex_if.get_method("semilagrangian").__module__.startswith("pysteps_b200")
        V = _quiet(mo_if.get_method("lk"), fr)
        assert np.abs(V - V_stock).max() <= 1e-12
        got = _quiet(steps.forecast, R, V_stock, **kw)
        got_resident = _quiet(steps.forecast, R, V_stock, extrap_kwargs={"b200_resident": True}, **kw)
    finally:
        for reg, old in zip((ex_if._extrapolation_methods, mo_if._methods, no_if._noise_methods), saved):
            reg.clear()
            reg.update(old)
    assert want.shape == got.shape == (3, 4, m, n) and np.isfinite(want).any()
    assert np.array_equal(want, got, equal_nan=True)
    assert np.array_equal(want, got_resident, equal_nan=True)


def test_nowcast_main_loop_ensemble_with_b200_methods(ref):
    import pysteps_b200
    from pysteps_b200 import _synthetic as syn
    utils = ref.ref_module("pysteps.nowcasts.utils")
    noise = ref.ref_module("pysteps.noise.interface")
    pysteps_b200.register()
    m, n, members = 200, 240, 3
    precip = syn.rain_field(m, n, 5)
    velocity = 2.0 * syn.velocity_field(m, n, 5)
    params = {"decay": 0.97, "bias": np.array([0.0, 0.1, -0.05])}
    state0 = np.stack([precip * (1 + 0.05 * i) for i in range(members)])

    def model(state, params):
        fields = state["fields"] * params["decay"] + params["bias"][:, None, None]
        return fields, {"fields": fields}

    def run(noise_name, extrap_name, extrap_kwargs, timesteps):
        init, gen = noise.get_method(noise_name)
        perts = []
        for j in range(members):
            vp = init(velocity, 1.0, 5.0, randstate=np.random.RandomState(100 + j))
            perts.append(lambda t, vp=vp: gen(vp, t * 5.0))   # nowcasts/steps.py:927-929
        return utils.nowcast_main_loop(precip, velocity, {"fields": state0.copy()}, timesteps, extrap_name, model,
                                       extrap_kwargs=extrap_kwargs, velocity_pert_gen=perts, params=params,
                                       ensemble=True, num_ensemble_members=members)

    for timesteps in (3, [0.5, 1.0, 2.25, 3.0]):
        want = _quiet(run, "bps", "semilagrangian", {"allow_nonfinite_values": True}, timesteps)
        got = _quiet(run, "bps_b200", "semilagrangian_b200", {"allow_nonfinite_values": True, "b200_resident": True},
                     timesteps)
        for g_member, w_member in zip(got, want):
            for g, w in zip(g_member, w_member):
                assert np.array_equal(g, w, equal_nan=True)
