"""Pins oracle/cost_volume.{c,py}: two independent restatements of
BuildCostVolume_cuda.cu:15-98 agree bit-for-bit, and both reproduce hand-computed cases and the
algebraic properties of SURVEY.md Appendix A."""
import numpy as np
import pytest

from oracle import cost_volume as cv


def _rand(shape, seed, dtype=np.float32):
    return np.random.default_rng(seed).standard_normal(shape).astype(dtype)


@pytest.mark.parametrize("ds", [1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c_and_numpy_restatements_bit_equal(ds, dtype):
    l, r = _rand((2, 3, 6, 10), 0, dtype), _rand((2, 3, 6, 10), 1, dtype)
    s = np.array([[0, 0.25, 1.0, 8.999, 9.0, 12.0, 1e-4], [0.5, 2.5, 3.75, 9.0, 8.5, 0.0, 100.0]], dtype=dtype)
    a = cv.forward_c(l, r, s, ds, fma_mode=0)
    b = cv.forward_np(l, r, s, ds)
    assert a.dtype == dtype and a.shape == (2, 6, 7, 6 // ds, 10 // ds)
    assert np.array_equal(a, b)
    c = cv.forward_c(l, r, s, ds, fma_mode=1)          # nvcc-style contraction: <= 1 ulp away
    assert np.max(np.abs(a - c)) <= 4 * np.finfo(dtype).eps * np.max(np.abs(a))


def test_hand_computed_row():
    # one row, W=4: R = [10, 20, 30, 40]; shift 0.25 -> x = w - 0.25
    l = np.arange(4, dtype=np.float32).reshape(1, 1, 1, 4)
    r = np.array([10, 20, 30, 40], dtype=np.float32).reshape(1, 1, 1, 4)
    s = np.array([[0.25, 1.0, 3.0, 3.5]], dtype=np.float32)
    out = cv.forward_c(l, r, s, 1)
    assert out.shape == (1, 2, 4, 1, 4)
    assert np.array_equal(out[0, 0, :, 0, :], np.tile(np.arange(4, dtype=np.float32), (4, 1)))
    np.testing.assert_array_equal(out[0, 1, 0, 0], np.float32([0, 17.5, 27.5, 37.5]))   # x=-.25 invalid
    np.testing.assert_array_equal(out[0, 1, 1, 0], np.float32([0, 10, 20, 30]))
    np.testing.assert_array_equal(out[0, 1, 2, 0], np.float32([0, 0, 0, 10]))
    np.testing.assert_array_equal(out[0, 1, 3, 0], np.float32([0, 0, 0, 0]))


def test_properties():
    N, C, H, W, D = 2, 4, 5, 12, 6
    l, r = _rand((N, C, H, W), 2), _rand((N, C, H, W), 3)
    zero = cv.forward_c(l, r, np.zeros((N, D), np.float32))
    assert np.array_equal(zero[:, C:], np.broadcast_to(r[:, :, None], (N, C, D, H, W)))   # shift 0 -> copy
    assert np.array_equal(zero[:, :C], np.broadcast_to(l[:, :, None], (N, C, D, H, W)))
    k = 3
    ints = cv.forward_c(l, r, np.full((N, D), float(k), np.float32))
    assert np.array_equal(ints[:, C:, :, :, k:], np.broadcast_to(r[:, :, None, :, :W - k], (N, C, D, H, W - k)))
    assert np.all(ints[:, C:, :, :, :k] == 0)
    far = cv.forward_c(l, r, np.full((N, D), W - 1 + 0.5, np.float32))
    assert np.all(far[:, C:] == 0)
    edge = cv.forward_c(l, r, np.full((N, D), 0.0, np.float32))      # x == W-1 is valid (clamped corner)
    assert np.array_equal(edge[:, C:, :, :, W - 1], np.broadcast_to(r[:, :, None, :, W - 1], (N, C, D, H)))


def test_negative_shift_rejected_and_empty():
    l = np.zeros((1, 1, 2, 2), np.float32)
    with pytest.raises(AssertionError):           # build_cost_volume/__init__.py:12
        cv.forward_c(l, l, np.float32([[-1.0]]))
    out = cv.forward_c(l[:0], l[:0], np.zeros((0, 3), np.float32))
    assert out.shape == (0, 2, 3, 2, 2)


def test_xlow_matches_forward_validity():
    s = np.float32([[0, 0.25, 1.0, 8.999, 9.0, 12.0, 1e-4]])
    xl = cv.xlow_c(s, 10)
    x = np.arange(10, dtype=np.float32)[None, None] - s[:, :, None]
    assert np.array_equal(xl >= 0, (x >= 0) & (x <= 9))


def test_backward_is_adjoint_of_forward():
    N, C, H, W, D, ds = 1, 2, 4, 8, 3, 1
    rng = np.random.default_rng(5)
    l, r = rng.standard_normal((N, C, H, W)), rng.standard_normal((N, C, H, W))
    s = np.array([[0.0, 1.3, 2.75]])
    g = rng.standard_normal((N, 2 * C, D, H, W))
    gl, gr = cv.backward_c(g, s, ds)
    # <fwd(l, r), g> == <l, gl> + <r, gr>  (the op is linear in (l, r))
    lhs = np.sum(cv.forward_c(l, r, s, ds, fma_mode=0) * g)
    rhs = np.sum(l * gl) + np.sum(r * gr)
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))
