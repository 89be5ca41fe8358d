"""
Drop-in for the reference's fill_alpha_harmonic.py: replaces every pixel whose opacity is
not 100% with a harmonic function of the opaque ones.

    python -m harmonic_interpolation_b200.fill_alpha_harmonic path/to/input_image path/to/output_image

Same functions (fill_alpha_harmonic, load_img2arr, arr2img, main), same exit behaviour
(usage + exit -1 on bad arguments, a missing input or an EXISTING output,
fill_alpha_harmonic.py:47-79); the solve goes through libhgrid.so.
"""
import numpy as np

from .recovery import solve_grid_linear, solve_grid_linear_simple  # noqa: F401  (same imports as the reference)


def fill_alpha_harmonic(arr, mask, **solver_kwargs):
    """fill_alpha_harmonic.py:6-25: pixels where `mask` is True keep their value (hard
    constraints); the rest solve the Laplace equation.  arr is (rows, cols[, K]); returns
    float64 (rows, cols, K).  The mask and the image go to the solver as arrays, not as one
    Python tuple per opaque pixel."""
    arr = np.asarray(arr)
    if len(arr.shape) == 2:
        arr = arr[..., np.newaxis]
    mask = np.asarray(mask, dtype=bool)
    return solve_grid_linear_simple(arr.shape[0], arr.shape[1], value_constraints=(mask, arr), **solver_kwargs)


def fill_rgba8(rgba, precision=None, rtol=None):
    """The CLI's array pipeline in one call (fill_alpha_harmonic.py:72-76): uint8 RGBA (rows, cols, 4)
    in, uint8 RGB out; keep-mask alpha == 255; pixels travel as bytes both ways and are scaled /
    rounded on the device exactly as `u8 / 255.` and `(255 x).round(0).clip(0, 255)` do."""
    from .solver import GridSolver, encode_flags
    rgba = np.asarray(rgba, dtype=np.uint8)
    assert rgba.ndim == 3 and rgba.shape[2] == 4
    rows, cols = rgba.shape[:2]
    keep = rgba[:, :, 3] == 255
    assert keep.any()
    s = GridSolver(rows, cols, 3, False, precision)
    s.set_pattern(encode_flags(rows, cols, hard=keep))
    out = s.solve_u8(np.ascontiguousarray(rgba[:, :, :3]), rtol=rtol)
    s.close()
    return out


def load_img2arr(path):
    """fill_alpha_harmonic.py:27-36: image -> float64 (rows, cols, RGBA) in [0, 1]."""
    from PIL import Image
    img = Image.open(path).convert('RGBA')
    arr = np.asarray(img, dtype=np.uint8) / 255.
    return arr.astype(float)


def arr2img(arr):
    """fill_alpha_harmonic.py:38-45: float image in [0, 1] -> PIL image (round half to even, clip)."""
    from PIL import Image
    assert arr.dtype in (np.float32, np.float64, float)
    return Image.fromarray(np.asarray((arr * 255).round(0).clip(0, 255), dtype=np.uint8))


def main(argv=None):
    import os
    import sys

    if argv is None:
        argv = sys.argv

    def usage():
        print('Usage:', argv[0], 'path/to/input_image path/to/output_image', file=sys.stderr)
        print('Replaces pixels whose opacity is not 100% with a harmonic function. 100% opaque pixels remain unchanged.', file=sys.stderr)
        print('Example:', argv[0], '"fill_alpha_harmonic-test/test.png" "fill_alpha_harmonic-test/out.png"', file=sys.stderr)
        sys.exit(-1)

    try:
        inpath, outpath = argv[1:]
    except ValueError:
        usage()

    if not os.path.isfile(inpath):
        print('ERROR: Input path does not exist.', file=sys.stderr)
        usage()
    if os.path.exists(outpath):
        print('ERROR: Output path exists; not clobbering.', file=sys.stderr)
        usage()

    print('Loading input image:', inpath)
    arr = load_img2arr(inpath)

    print('Filling in non-opaque values...')
    result = fill_alpha_harmonic(arr[:, :, :3], (255 * arr[:, :, 3]).round(0).astype(int) == 255)
    result_img = arr2img(result)

    result_img.save(outpath)
    print('Saved output to:', outpath)


if __name__ == '__main__':
    main()
