"""Pins oracle/lsap_ref.c against the installed scipy (the function the reference calls,
losses_and_metrics.py:242) bit-for-bit, ties / +inf / tall / empty cases included."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from scipy.optimize import linear_sum_assignment

from util import lsap_c, lsap_c_solve


def _check(c):
    c = np.asarray(c, np.float32)
    try:
        r, cc = linear_sum_assignment(c)
        err = None
    except ValueError as e:
        err = str(e)
    k, a, b = lsap_c_solve(c)
    if err is None:
        assert k == len(r)
        assert (a == r).all() and (b == cc).all()
    else:
        assert k == (-2 if "invalid" in err else -1)


@pytest.mark.parametrize("kind", ["uniform", "small_int", "identical_cols", "inf", "eighths", "const"])
def test_random_cases(kind):
    rng = np.random.default_rng(hash(kind) % 2**32)
    for _ in range(400):
        nr, nc = int(rng.integers(0, 16)), int(rng.integers(1, 16))
        if kind == "uniform":
            c = rng.random((nr, nc))
        elif kind == "small_int":
            c = rng.integers(0, 3, (nr, nc)).astype(float)
        elif kind == "identical_cols":
            c = np.tile(rng.random((nr, 1)), (1, nc))
        elif kind == "inf":
            c = rng.integers(0, 4, (nr, nc)).astype(float)
            c[rng.random((nr, nc)) < 0.25] = np.inf
        elif kind == "eighths":
            c = np.round(rng.random((nr, nc)) * 8) / 8
        else:
            c = np.full((nr, nc), 3.25)
        _check(c)


def test_invalid_entries():
    c = np.ones((3, 4), np.float32)
    c[1, 2] = np.nan
    _check(c)
    c[1, 2] = -np.inf
    _check(c)
    _check(np.full((3, 3), np.inf))


@pytest.mark.parametrize("shape", [(20, 100), (100, 300), (100, 100), (120, 100), (1, 300), (300, 1)])
def test_realistic_sizes(shape):
    rng = np.random.default_rng(7)
    _check(rng.random(shape))
    _check(np.round(rng.random(shape) * 16) / 16)


def test_documented_behaviours():
    # constant matrix -> identity; tall -> sorted rows
    k, a, b = lsap_c_solve(np.full((4, 6), 2.0))
    assert list(b) == [0, 1, 2, 3]
    k, a, b = lsap_c_solve(np.random.default_rng(0).random((5, 3)))
    assert k == 3 and list(a) == sorted(a)
    assert lsap_c_solve(np.zeros((0, 5)))[0] == 0


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 9), st.integers(1, 9), st.integers(1, 4), st.randoms(use_true_random=False))
def test_hypothesis_ties(nr, nc, levels, rnd):
    c = np.array([[rnd.randint(0, levels) / 4.0 for _ in range(nc)] for _ in range(nr)], np.float32)
    _check(c)


def test_batch_mask_matches_reference_loop():
    import ctypes
    from oracle.reference_path import matching_assignment
    rng = np.random.default_rng(3)
    B, T, Q = 6, 7, 11
    cost = (np.round(rng.random((B, T, Q)) * 8) / 8).astype(np.float32)
    n = rng.integers(0, T + 1, B).astype(np.int32)
    ref = matching_assignment(cost, n)
    mask = np.empty_like(cost)
    rc = lsap_c().lsap_ref_batch_mask(cost.ctypes.data_as(ctypes.c_void_p), n.ctypes.data_as(ctypes.c_void_p),
                                      B, T, Q, mask.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0 and (mask == ref).all()
