"""Registration into pysteps' method registries (the drop-in boundary).

pysteps has no entry-point discovery for motion / extrapolation methods; the
"plugin API" is the module-level dict read by ``get_method``:
  pysteps/extrapolation/interface.py:107-111  ``_extrapolation_methods``
  pysteps/motion/interface.py:36-46           ``_methods``
  pysteps/noise/interface.py:24-45            ``_noise_methods``  ("bps": the velocity perturbator)
``register()`` inserts the B200 callables under new names and, on request,
under the stock names so that ``nowcasts.steps`` (which fetches the
extrapolator by name at pysteps/nowcasts/steps.py:656 and
pysteps/nowcasts/utils.py:359) runs unchanged.
"""


def methods():
    """name -> callable for everything this package provides."""
    from .extrapolation import semilagrangian

    from .noise import motion as bps

    out = {"extrapolation": {"semilagrangian_b200": semilagrangian.extrapolate}, "motion": {},
           "noise": {"bps_b200": (bps.initialize_bps, bps.generate_bps)}}
    try:
        from .motion import lucaskanade
        out["motion"]["lk_b200"] = lucaskanade.dense_lucaskanade
        out["motion"]["lucaskanade_b200"] = lucaskanade.dense_lucaskanade
    except ImportError:
        pass
    try:
        from .motion import vet
        out["motion"]["vet_b200"] = vet.vet
    except ImportError:
        pass
    from .motion import proesmans
    out["motion"]["proesmans_b200"] = proesmans.proesmans
    return out


def register(override=False):
    """Insert the B200 methods into an importable ``pysteps``.

    override=False: only the ``*_b200`` names are added (the identity checks of
    pysteps/tests/test_interfaces.py keep passing).  override=True additionally
    replaces ``"semilagrangian"``, ``"lk"``/``"lucaskanade"``, ``"vet"`` and the noise
    method ``"bps"``.
    Returns the list of registered names.
    """
    import pysteps.extrapolation.interface as ei
    import pysteps.motion.interface as mi
    import pysteps.noise.interface as ni

    done = []
    m = methods()
    for name, fn in m["extrapolation"].items():
        ei._extrapolation_methods[name] = fn
        done.append("extrapolation:" + name)
        if override:
            ei._extrapolation_methods[name.replace("_b200", "")] = fn
            done.append("extrapolation:" + name.replace("_b200", ""))
    for name, fn in m["motion"].items():
        mi._methods[name] = fn
        done.append("motion:" + name)
        if override:
            mi._methods[name.replace("_b200", "")] = fn
            done.append("motion:" + name.replace("_b200", ""))
    for name, fns in m["noise"].items():
        ni._noise_methods[name] = fns
        done.append("noise:" + name)
        if override:
            ni._noise_methods[name.replace("_b200", "")] = fns
            done.append("noise:" + name.replace("_b200", ""))
    return done
