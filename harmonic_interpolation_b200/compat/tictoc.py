"""Top-level name `tictoc` for callers of the reference (`from tictoc import ...`): put this directory
on PYTHONPATH.  Everything is re-exported from harmonic_interpolation_b200.tictoc."""
from harmonic_interpolation_b200.tictoc import *  # noqa: F401,F403
from harmonic_interpolation_b200 import tictoc as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
