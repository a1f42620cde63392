"""Import the *unmodified* Python reference from /root/reference -- TEST INFRASTRUCTURE ONLY.

Used by tests/golden/make_golden.py (and optional cross-checks) in the build
container; /root/reference does not exist on the GPU box, so nothing that runs
there may depend on this module.

Shims (none touches reference source):
  * matplotlib is not installed -> stub modules injected into sys.modules;
  * /root/reference is read-only -> private NUMBA_CACHE_DIR (one per variant);
  * variant "toric": planar_model._apply_random_stabilizer is rebound to the
    toric proposal function *before* src.mcmc is imported, so the fast chain
    kernel uses toric geometry.  Variant "shipped" leaves it alone.
One variant per process (numba caches compiled callees).
"""
import os
import sys
import types

REF = os.environ.get("QEC_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "src"))


def load(variant="shipped"):
    assert variant in ("shipped", "toric")
    assert available(), "reference checkout not present"
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/qec_numba_cache_" + variant)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(show=lambda *a, **k: None, figure=lambda *a, **k: None,
                              subplots=lambda *a, **k: (None, None), savefig=lambda *a, **k: None,
                              close=lambda *a, **k: None)
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.toric_model as toric_model
    import src.planar_model as planar_model
    if variant == "toric":
        planar_model._apply_random_stabilizer = toric_model._apply_random_stabilizer
    import src.rotated_surface_model as rotated_model
    import src.xzzx_model as xzzx_model
    import src.mcmc as mcmc
    import src.mcmc_alpha as mcmc_alpha
    import src.mcmc_biased as mcmc_biased
    import decoders
    import decoders_biasednoise
    ns = types.SimpleNamespace(
        toric_model=toric_model, planar_model=planar_model, rotated_model=rotated_model,
        xzzx_model=xzzx_model, mcmc=mcmc, mcmc_alpha=mcmc_alpha, mcmc_biased=mcmc_biased,
        decoders=decoders, decoders_biasednoise=decoders_biasednoise, variant=variant)
    ns.codes = {"toric": toric_model.Toric_code, "planar": planar_model.Planar_code,
                "rotated": rotated_model.RotSurCode, "xzzx": xzzx_model.xzzx_code}
    ns.models = {"toric": toric_model, "planar": planar_model, "rotated": rotated_model, "xzzx": xzzx_model}

    import random as pyrandom
    import numpy as np
    from numba import njit

    @njit
    def _nb_seed(s):
        pyrandom.seed(s)

    def seed_all(py=None, nb=None, np_=None):
        """Seed the reference's three independent RNG streams."""
        if py is not None:
            pyrandom.seed(py)
        if nb is not None:
            _nb_seed(nb)
        if np_ is not None:
            np.random.seed(np_)

    ns.seed_all = seed_all
    return ns
