"""Import alias: the package sources live in ``mcmc-qec-toric-rl_b200/`` (not a valid
Python identifier), this stub makes them importable as ``mcmc_qec_toric_rl_b200``."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mcmc-qec-toric-rl_b200"))
from ._version import __version__  # noqa: E402,F401
