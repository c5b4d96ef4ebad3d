"""Seeded semi-Lagrangian test cases shared by the golden generator and the tests."""
import numpy as np

from pysteps_b200 import _synthetic as syn

M, N = 56, 72


def _base(seed=0, kind="smooth"):
    P = syn.rain_field(M, N, seed)
    V = syn.velocity_field(M, N, seed, kind)
    return P, V


def build_case(name):
    """-> (args, kwargs) for extrapolate(precip, velocity, timesteps, **kwargs)."""
    P, V = _base(0)
    if name == "default_T4":
        return (P, V, 4), {}
    if name == "rotation_T5":
        P, V = _base(1, "rotation")
        return (P, V * 4.0, 5), {}
    if name == "list_timesteps":
        return (P, V, [0.5, 1.0, 2.5]), {"vel_timestep": 2.0}
    if name == "n_iter3":
        return (P, V, 3), {"n_iter": 3}
    if name == "n_iter0":
        return (P, V, 3), {"n_iter": 0}
    if name == "n_iter0_prev":
        D = 0.25 * syn.velocity_field(M, N, 7)
        return (P, V, 2), {"n_iter": 0, "displacement_prev": D, "return_displacement": True}
    if name == "nearest_mode":
        return (P, V * 3.0, 3), {"map_coordinates_mode": "nearest"}
    if name == "outval_min":
        return (P - 3.0, V * 3.0, 3, "min"), {}
    if name == "outval_const":
        return (P, V * 3.0, 2, -15.0), {}
    if name == "return_disp":
        return (P, V, 3), {"return_displacement": True}
    if name == "disp_prev":
        D = -1.5 * syn.velocity_field(M, N, 3)
        return (P, V, [1.0]), {"displacement_prev": D, "return_displacement": True}
    if name == "precip_none":
        return (None, V, 3), {"return_displacement": True}
    if name == "nan_disc":
        return (syn.nan_disc(P), V, 3), {"allow_nonfinite_values": True}
    if name == "float32_precip":
        return (P.astype(np.float32), V, 3), {}
    if name == "float32_both":
        return (P.astype(np.float32), V.astype(np.float32), 3), {"return_displacement": True}
    if name == "custom_xy":
        x, y = np.meshgrid(np.arange(N) * 0.75 + 2.0, np.arange(M) * 0.9 + 1.5)
        return (P, V, 2), {"xy_coords": np.stack([x, y])}
    if name == "default_xy_given":
        x, y = np.meshgrid(np.arange(N), np.arange(M))
        return (P, V, 2), {"xy_coords": np.stack([x, y])}
    if name == "long_T40":
        return (P, V * 0.2, 40), {"return_displacement": True}
    if name == "noisy_velocity":
        rng = np.random.default_rng(11)
        return (P, V + rng.normal(size=V.shape), 3), {}
    # ---- spline orders (semilagrangian.py:144-157,224-253); oracle only so far ----------------
    if name == "order3_constant":
        return (P, V * 2.0, 3), {"interp_order": 3}
    if name == "order3_nearest_nan":
        # the call of examples/ens_kalman_filter_blended_forecast.py:258 (order 3, "nearest")
        return (syn.nan_disc(P), V * 3.0, 3), {"interp_order": 3, "map_coordinates_mode": "nearest",
                                               "allow_nonfinite_values": True}
    if name == "order3_float32_min":
        return (P.astype(np.float32) - 2.0, V * 3.0, 2, "min"), {"interp_order": 3, "return_displacement": True}
    if name == "order0_nearest":
        return (P, V * 3.0, 2), {"interp_order": 0, "map_coordinates_mode": "nearest"}
    if name == "order0_constant":
        return (P, V * 3.0, 2, -15.0), {"interp_order": 0}
    raise KeyError(name)


# cases the CUDA path implements (interp_order 1)
CASES = ["default_T4", "rotation_T5", "list_timesteps", "n_iter3", "n_iter0", "n_iter0_prev",
         "nearest_mode", "outval_min", "outval_const", "return_disp", "disp_prev", "precip_none",
         "nan_disc", "float32_precip", "float32_both", "custom_xy", "default_xy_given",
         "long_T40", "noisy_velocity"]

# cases only the oracle restates so far (the CUDA path raises NotImplementedError for them)
SPLINE_CASES = ["order3_constant", "order3_nearest_nan", "order3_float32_min", "order0_nearest",
                "order0_constant"]
