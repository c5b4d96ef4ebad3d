"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy) of the BPS motion perturbator,
pysteps/noise/motion.py:55-180, and of the call-site expression
pysteps/nowcasts/utils.py:448-451 ``velocity + velocity_pert_gen[i](t)``.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the
product (pysteps_b200/) never does.

Pinned: tests/test_oracle_bps.py compares it bit for bit with the reference's outputs stored
in tests/golden/bps_golden.npz (made by tests/golden/gen_bps_golden.py from the reference
imported in the build container) and, when /root/reference is present, with the reference
itself.
"""
import numpy as np

DEFAULT_PAR = (10.88, 0.23, -7.68)   # motion.py:43-46
DEFAULT_PERP = (5.76, 0.31, -2.72)   # motion.py:49-52


def unit_vectors(V):
    """motion.py:134-139: V/|V| as a float64 (2,m,n) array, zero where |V| <= 1e-12.  The norm
    and the quotient are evaluated in V's own floating dtype (scipy.linalg.norm ->
    numpy.linalg.norm -> sqrt(add.reduce(x*x, axis=0)); integer input is promoted to float64);
    non-finite input raises like scipy.linalg.norm(check_finite=True)."""
    V = np.asarray(V)
    if not np.issubdtype(V.dtype, np.inexact):
        V = V.astype(np.float64)
    if not np.all(np.isfinite(V)):
        raise ValueError("array must not contain infs or NaNs")
    speed = np.sqrt(V[0] * V[0] + V[1] * V[1])
    moving = speed > 1e-12
    unit = np.zeros(V.shape, dtype=np.float64)
    for c in range(2):
        q = np.zeros(speed.shape, dtype=V.dtype)
        np.divide(V[c], speed, out=q, where=moving)
        unit[c] = q
    return unit


def initialize_bps(V, pixelsperkm, timestep, p_par=None, p_perp=None, randstate=None, seed=None):
    """motion.py:55-141"""
    if len(np.shape(V)) != 3:
        raise ValueError("V is not a three-dimensional array")
    if np.shape(V)[0] != 2:
        raise ValueError("the first dimension of V is not 2")
    p_par = DEFAULT_PAR if p_par is None else p_par
    p_perp = DEFAULT_PERP if p_perp is None else p_perp
    if len(p_par) != 3:
        raise ValueError("the length of p_par is not 3")
    if len(p_perp) != 3:
        raise ValueError("the length of p_perp is not 3")
    rs = np.random if randstate is None else randstate
    if seed is not None:
        rs.seed(seed)
    eps_par = rs.laplace(scale=1.0 / np.sqrt(2))
    eps_perp = rs.laplace(scale=1.0 / np.sqrt(2))
    unit = unit_vectors(V)
    return {"randstate": rs, "vsf": 60.0 / (timestep * pixelsperkm), "p_par": p_par, "p_perp": p_perp,
            "eps_par": eps_par, "eps_perp": eps_perp, "V_par": unit,
            "V_perp": np.stack([-unit[1], unit[0]])}


def coefficients(perturbator, t):
    """(a_par, a_perp) = (g_par(t)*eps_par, g_perp(t)*eps_perp), motion.py:172-180"""
    pp, pq = perturbator["p_par"], perturbator["p_perp"]
    g_par = pp[0] * pow(t, pp[1]) + pp[2]
    g_perp = pq[0] * pow(t, pq[1]) + pq[2]
    return g_par * perturbator["eps_par"], g_perp * perturbator["eps_perp"]


def generate_bps(perturbator, t):
    """motion.py:144-180"""
    a, b = coefficients(perturbator, t)
    return (a * perturbator["V_par"] + b * perturbator["V_perp"]) / perturbator["vsf"]


def perturbed_velocity(V, perturbator, t):
    """nowcasts/utils.py:448-451"""
    return V + generate_bps(perturbator, t)
