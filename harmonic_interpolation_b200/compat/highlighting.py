"""Top-level name `highlighting` for callers of the reference (`from highlighting import ...`): put this directory
on PYTHONPATH.  Everything is re-exported from harmonic_interpolation_b200.highlighting."""
from harmonic_interpolation_b200.highlighting import *  # noqa: F401,F403
from harmonic_interpolation_b200 import highlighting as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
