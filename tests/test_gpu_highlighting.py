"""GPU: the two tint functions of highlighting.py against outputs of the unmodified reference."""
import contextlib
import io

import numpy as np
import pytest

import highlighting_cases as hc
from harmonic_interpolation_b200 import highlighting as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", hc.NAMES)
def test_tint_functions_match_the_reference(name):
    case = hc.load(name)
    with contextlib.redirect_stdout(io.StringIO()):
        got = hc.run(H, case)
    want = case["out"]
    assert got.dtype == np.uint8 and got.shape == want.shape
    # the input image is untouched and only the regions change
    assert (case["img"] == hc.load(name)["img"]).all()
    diff = np.abs(got.astype(int) - want.astype(int))
    # the blend is rounded to uint8: a solve that agrees to 1e-9 can still flip a .5 case by one level
    assert diff.max() <= 1, diff.max()
    assert (diff > 0).mean() < 1e-3
