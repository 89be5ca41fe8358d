"""Nested stopwatch printing in the style of the reference's tictoc.py:5-20
('+' * depth on entry, '-' * depth, message and seconds on exit), on perf_counter
(the reference's time.clock no longer exists)."""
import time
from contextlib import contextmanager

starts = []
quiet = False


def tic(msg=None):
    if msg is not None and not quiet:
        print("+" * (len(starts) + 1), msg)
    starts.append((time.perf_counter(), msg))


def toc():
    t0, msg = starts.pop()
    duration = time.perf_counter() - t0
    if not quiet:
        if msg is None:
            print("=" * (len(starts) + 1), "tictoc():", duration)
        else:
            print("-" * (len(starts) + 1), msg, duration)
    return duration


@contextmanager
def tictoc(msg=None):
    tic(msg)
    try:
        yield
    finally:
        toc()


def tictoc_dec(func):
    """tictoc.py:28-35: decorator that wraps every call of `func` in tictoc(func.__name__)."""
    def wrapped(*args, **kwargs):
        with tictoc(func.__name__):
            return func(*args, **kwargs)
    return wrapped
