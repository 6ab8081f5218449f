"""Oracle tableaus against independent anchors: SciPy's Dormand-Prince (RK45.{A,B,C,E,P}) and the
Runge-Kutta order conditions (Tsit5 has no copy in this image)."""

import numpy as np
import pytest
from scipy.integrate._ivp.rk import RK45

from oracle import tableaus as T


def _dense(tab):
    s = tab.n_stages
    A = np.zeros((s, s))
    for i, row in enumerate(tab.a):
        A[i, :len(row)] = row
    return A, np.array(tab.b), np.array(tab.c)


def test_dopri5_matches_scipy():
    A, b, c = _dense(T.DOPRI5)
    assert np.allclose(A[:6, :5], RK45.A[:, :5], rtol=0, atol=1e-16)
    assert np.allclose(b[:6], RK45.B, rtol=0, atol=1e-16)
    assert np.allclose(c[:6], RK45.C, rtol=0, atol=1e-16)
    assert np.allclose(A[6, :6], RK45.B, rtol=0, atol=1e-16)          # FSAL row
    assert np.allclose(np.abs(T.DOPRI5.e), np.abs(RK45.E), rtol=0, atol=1e-16)
    mid = RK45.P @ np.array([0.5, 0.25, 0.125, 0.0625])              # dense output at theta = 1/2
    assert np.allclose(T.DOPRI5.b_mid, mid, rtol=0, atol=1e-15)


def _order_conditions(A, b, c, order):
    res = [b.sum() - 1]
    if order >= 2:
        res += [b @ c - 1 / 2]
    if order >= 3:
        res += [b @ c**2 - 1 / 3, b @ A @ c - 1 / 6]
    if order >= 4:
        res += [b @ c**3 - 1 / 4, (b * c) @ A @ c - 1 / 8, b @ A @ c**2 - 1 / 12, b @ A @ A @ c - 1 / 24]
    if order >= 5:
        res += [b @ c**4 - 1 / 5, (b * c**2) @ A @ c - 1 / 10, (b * c) @ A @ c**2 - 1 / 15,
                (b * c) @ A @ A @ c - 1 / 30, b @ (A @ c) ** 2 - 1 / 20, b @ A @ c**3 - 1 / 20,
                b @ A @ (c * (A @ c)) - 1 / 40, b @ A @ A @ c**2 - 1 / 60, b @ A @ A @ A @ c - 1 / 120]
    return np.abs(np.array(res)).max()


@pytest.mark.parametrize("name,order,low", [("dopri5", 5, 4), ("tsit5", 5, 4), ("heun", 2, 1),
                                            ("euler", 1, None), ("rk4", 4, None), ("rk4_38", 4, None)])
def test_order_conditions(name, order, low):
    tab = T.BY_NAME[name]
    A, b, c = _dense(tab)
    assert np.abs(A.sum(1) - c).max() < 1e-15
    assert _order_conditions(A, b, c, order) < 5e-15
    assert tab.order == order
    if low is not None:                      # embedded method b_low = b - e is of order `low`
        assert _order_conditions(A, b - np.array(tab.e), c, low) < 5e-15
        assert _order_conditions(A, b - np.array(tab.e), c, low + 1) > 1e-6


def test_tsit5_dense_output_consistency():
    assert np.allclose(T.tsit5_dense_weights(1.0), T.TSIT5.b, rtol=0, atol=1e-14)
    assert np.allclose(T.tsit5_dense_weights(0.0), 0.0)
    assert np.allclose(T.TSIT5.b_mid, T.tsit5_dense_weights(0.5))
    assert abs(sum(T.TSIT5.b_mid) - 0.5) < 1e-14              # interpolant reproduces y' = 1


def test_fsal_rows():
    for tab in (T.DOPRI5, T.TSIT5):
        assert tab.fsal and tab.ssal
        assert np.allclose(tab.a[-1], tab.b[:-1]) and tab.b[-1] == 0.0
