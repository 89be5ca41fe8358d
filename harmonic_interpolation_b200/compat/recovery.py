"""Top-level name `recovery` for callers of the reference (`from recovery import ...`): put this directory
on PYTHONPATH.  Everything is re-exported from harmonic_interpolation_b200.recovery."""
from harmonic_interpolation_b200.recovery import *  # noqa: F401,F403
from harmonic_interpolation_b200 import recovery as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
