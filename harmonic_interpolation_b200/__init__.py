"""
harmonic_interpolation_b200 -- B200-native (sm_100a) grid-diffusion solve path behind the
call signatures of yig/harmonic_interpolation.

  recovery              drop-in for the reference's `recovery` solve entry points
  fill_alpha_harmonic   drop-in for the reference's CLI / array function
  highlighting          `grow` / `shrink` (used by smooth_bumps*)
  solver.GridSolver     array-native interface to libhgrid.so (include/hgrid.h)

Every numerical path goes through hand-written CUDA kernels in libhgrid.so; there is no
CPU fallback: importing works anywhere (so host logic can be tested), but creating a
solver without the library or without a GPU raises.
"""
__version__ = "0.1.0"
