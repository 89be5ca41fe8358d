"""Top-level name `poisson` for callers of the reference (`from poisson import solve_poisson_simple`): put this
directory on PYTHONPATH.  Everything is re-exported from harmonic_interpolation_b200.poisson."""
from harmonic_interpolation_b200.poisson import *  # noqa: F401,F403
from harmonic_interpolation_b200 import poisson as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
