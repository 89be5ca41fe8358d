"""`grow` / `shrink` with the behaviour of the reference's highlighting.py:26-74:
L1-ball (diamond) dilation of a boolean mask, one 4-neighbour OR per iteration."""
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
