"""The in-tree callers of the grid solve in the reference's highlighting.py: `grow` / `shrink` (highlighting.py:26-74:
L1-ball (diamond) dilation of a boolean mask, one 4-neighbour OR per iteration) and the two tint functions that solve a
small cropped (bi)laplacian problem per region (highlighting.py:139-213, 361-464).  The rest of that file is UI drawing
(PIL outlines, desaturation) and is out of scope."""
import numpy as np


def grow(mask, pixel_iterations):
    assert int(pixel_iterations) == pixel_iterations
    assert pixel_iterations >= 0
    mask = np.array(mask, dtype=bool)
    for _ in range(int(pixel_iterations)):
        prev = mask.copy()
        mask[:-1, :] |= prev[1:, :]
        mask[1:, :] |= prev[:-1, :]
        mask[:, :-1] |= prev[:, 1:]
        mask[:, 1:] |= prev[:, :-1]
    return mask


def shrink(mask, pixel_iterations):
    return np.logical_not(grow(np.logical_not(mask), pixel_iterations))


def _region_mask(shape, region):
    m = np.zeros(shape, dtype=bool)
    m[region] = True
    return m


def array_tinting_region_with_falloff_from_center(img, region, color, alpha, *, precision=None):
    """highlighting.py:139-213.  A copy of the uint8 rgb image `img` with `img[region]` tinted by `color`: strongest at the
    centre of the region, falling off towards its boundary -- the profile is a thin-plate bump, the solution of a
    bilaplacian problem on a crop around the region with the region's rim held at 1 and the ring just outside at 0.

    (The reference computes the rim as `r - rinner` on boolean arrays, highlighting.py:184, which numpy no longer accepts;
    this is the same set, r AND NOT rinner.)"""
    from . import recovery
    assert img.dtype == np.uint8
    assert len(img.shape) == 3
    assert img.shape[2] == 3
    assert len(color) == 3
    alpha = max(0., min(1., float(alpha)))

    # only a crop around the region is solved, with one buffer pixel so that the region has a proper boundary even at the image edge
    region_mask = _region_mask(img.shape[:2], region)
    kBufferPixels = 1
    ii, jj = np.nonzero(region_mask)
    crop = ii.min(), ii.max() + 1, jj.min(), jj.max() + 1
    r = np.zeros((crop[1] - crop[0] + 2 * kBufferPixels, crop[3] - crop[2] + 2 * kBufferPixels), dtype=bool)
    r[kBufferPixels:-kBufferPixels, kBufferPixels:-kBufferPixels][region_mask[crop[0]:crop[1], crop[2]:crop[3]]] = True

    router = grow(r, 1)
    router[r] = False
    rinner = np.logical_not(grow(np.logical_not(r), 1))
    rim = r & ~rinner
    values = np.zeros(r.shape)
    values[rim] = 1.
    print('len( value_constraints ):', int(rim.sum() + router.sum()))
    bump = recovery.solve_grid_linear(r.shape[0], r.shape[1], bilaplacian=True, value_constraints=(rim | router, values),
                                      iterative=False, precision=precision)
    bump[np.logical_not(r)] = 0.
    assert bump.min() >= -1e-5
    bump *= 1. / bump.max()
    bump = bump.clip(0., 1.)

    result = np.array(img)
    bump *= alpha
    color = np.asarray(color)
    result_slice = result[crop[0]:crop[1], crop[2]:crop[3]]
    result_slice[:] = (result_slice + bump[kBufferPixels:-kBufferPixels, kBufferPixels:-kBufferPixels, np.newaxis]
                       * (color - result_slice)).round().astype(np.uint8).clip(0, 255)
    return result


def array_tinting_region_border_with_falloff(img, regions, colors, alpha, *, precision=None):
    """highlighting.py:361-464.  A copy of `img` in which each of the abutting `regions` is tinted by its colour most
    strongly along the border it shares with the other regions, falling off to nothing at the rest of its boundary: per
    region one harmonic (Laplacian) problem on a crop with a two-pixel buffer -- 0 outside the region, 1 where the region
    grown by a pixel meets another grown region."""
    from . import recovery
    assert img.dtype == np.uint8
    assert len(img.shape) == 3
    assert img.shape[2] == 3
    assert len(set([id(r) for r in regions])) == len(regions)
    assert False not in [len(color) == 3 for color in colors]
    alpha = max(0., min(1., float(alpha)))

    result = np.array(img)
    masks = [_region_mask(img.shape[:2], region) for region in regions]
    grown_masks = [grow(mask, 1) for mask in masks]

    kBufferPixels = 2
    for idx, (mask, grown_mask, color) in enumerate(zip(masks, grown_masks, colors)):
        ii, jj = np.nonzero(mask)
        crop = ii.min(), ii.max() + 1, jj.min(), jj.max() + 1
        shape = (crop[1] - crop[0] + 2 * kBufferPixels, crop[3] - crop[2] + 2 * kBufferPixels)
        crop_mask = np.zeros(shape, dtype=bool)
        crop_mask[kBufferPixels:-kBufferPixels, kBufferPixels:-kBufferPixels][mask[crop[0]:crop[1], crop[2]:crop[3]]] = True

        # the ones: this region grown AND another region grown, pasted into the canvas (pixels that fall off the canvas are
        # dropped: the reference would index out of range there, which its two-pixel buffer makes impossible for abutting regions)
        ones_img = np.zeros(img.shape[:2], dtype=bool)
        for other, other_grown in enumerate(grown_masks):
            if other != idx:
                ones_img |= grown_mask & other_grown
        oi, oj = np.nonzero(ones_img)
        oi = oi - crop[0] + kBufferPixels
        oj = oj - crop[2] + kBufferPixels
        one = np.zeros(shape, dtype=bool)
        one[oi, oj] = True
        zero = np.logical_not(crop_mask)
        # (highlighting.py:431: the two sets are meant to overlap; the ones win)
        assert (one & zero).any()
        zero &= ~one

        values = np.zeros(shape)
        values[one] = 1.
        blend = recovery.solve_grid_linear(shape[0], shape[1], value_constraints=(one | zero, values), bilaplacian=False, precision=precision)
        blend[np.logical_not(crop_mask)] = 0.
        blend *= alpha

        color = np.asarray(color)
        result_slice = result[crop[0]:crop[1], crop[2]:crop[3]]
        result_slice[:] = (result_slice + blend[kBufferPixels:-kBufferPixels, kBufferPixels:-kBufferPixels, np.newaxis]
                           * (color - result_slice)).round().astype(np.uint8).clip(0, 255)
    return result
