"""CPU anchors of the CDE oracle (oracle/torchcde_like.py, oracle/torchdiffeq_like.py, oracle/pose_cde.py):
nothing here can be checked against the real torchcde / torchdiffeq (not installable offline), so the
restatement is pinned on what IS available: SciPy's converged solutions and Dormand-Prince dense
output, analytic convergence orders, and the interpolation conditions of the control paths."""

import math

import numpy as np
import pytest
import torch
from scipy.integrate import solve_ivp

from oracle import torchcde_like as cde
from oracle.torchdiffeq_like import NEXT, NONE, PREV, odeint_dopri5, odeint_rk4


def test_rectilinear_coeffs_layout():
    """(t1,x1),(t2,x1),(t2,x2),...: time moves first, then the values."""
    x = torch.tensor([[[0.1, 10.0], [0.3, 20.0], [0.6, 30.0]]])
    c = cde.linear_interpolation_coeffs(x, rectilinear=0)
    assert c.shape == (1, 5, 2)
    assert torch.equal(c[0], torch.tensor([[0.1, 10.0], [0.3, 10.0], [0.3, 20.0], [0.6, 20.0], [0.6, 30.0]]))
    X = cde.LinearInterpolation(c)
    assert torch.equal(X.grid_points, torch.arange(5.0))
    assert torch.equal(X.evaluate(X.interval[0]), x[:, 0])
    # even segments move only time, odd segments only the values
    assert torch.allclose(X.derivative(torch.tensor(0.5)), torch.tensor([[0.2, 0.0]]))
    assert torch.allclose(X.derivative(torch.tensor(1.5)), torch.tensor([[0.0, 10.0]]))
    # a t exactly on knot k > 0 belongs to segment k - 1; one ulp later to segment k
    one = torch.tensor(1.0)
    assert torch.allclose(X.derivative(one), torch.tensor([[0.2, 0.0]]))
    assert torch.allclose(X.derivative(torch.nextafter(one, one + 1)), torch.tensor([[0.0, 10.0]]))
    # beyond the grid: clamped to the last segment
    assert torch.allclose(X.derivative(torch.tensor(99.0)), torch.tensor([[0.0, 10.0]]))


def test_hermite_cubic_interpolation_conditions():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 3, generator=g, dtype=torch.float64)
    X = cde.HermiteCubicBackward(x)
    eps = 1e-9
    for i in range(5):
        a, b = torch.tensor(float(i) + eps, dtype=torch.float64), torch.tensor(float(i + 1), dtype=torch.float64)
        assert torch.allclose(X.evaluate(a), x[:, i], atol=1e-7)            # X(knot i) = x_i
        assert torch.allclose(X.evaluate(b), x[:, i + 1], atol=1e-12)       # X(knot i+1) = x_{i+1}
        d = x[:, i + 1] - x[:, i]
        m = d if i == 0 else x[:, i] - x[:, i - 1]
        assert torch.allclose(X.derivative(a), m, atol=1e-7)                # X'(knot i+) = backward difference
        assert torch.allclose(X.derivative(b), d, atol=1e-12)               # X'(knot i+1-) = this segment's slope
        # derivative is the derivative of evaluate
        t = torch.tensor(i + 0.37, dtype=torch.float64)
        h = 1e-6
        fd = (X.evaluate(t + h) - X.evaluate(t - h)) / (2 * h)
        assert torch.allclose(X.derivative(t), fd, atol=1e-7)


def _lin_system(dtype=torch.float64):
    A = torch.tensor([[-0.5, 2.0], [-2.0, -0.5]], dtype=dtype)
    return A, (lambda t, y, perturb=NONE: y @ A.T)


def test_dopri5_matches_scipy_solution_and_dense_output():
    A, f = _lin_system()
    y0 = torch.tensor([[1.0, 0.5], [-0.3, 0.8]], dtype=torch.float64)
    t = [0.0, 0.4, 1.1, 2.0]
    sol = odeint_dopri5(f, y0, t, rtol=1e-8, atol=1e-10)
    for b in range(2):
        ref = solve_ivp(lambda tt, y: A.numpy() @ y, (0.0, 2.0), y0[b].numpy(), method="RK45", rtol=1e-11,
                        atol=1e-13, t_eval=t)
        assert np.allclose(sol["ys"][:, b].numpy(), ref.y.T, rtol=0, atol=5e-8)
    assert sol["n_accepted"] <= sol["n_steps"] and sol["n_f_evals"] == 2 + 6 * sol["n_steps"]


def test_dopri5_controller_rules():
    """accept iff ratio <= 1; dt_next = dt * min(10, max(0.9 ratio^-0.2, dfactor)); first step by Hairer."""
    A, f = _lin_system()
    y0 = torch.tensor([[1.0, 0.5]], dtype=torch.float64)
    sol = odeint_dopri5(f, y0, [0.0, 3.0], rtol=1e-4, atol=1e-6)
    dts, ratios = sol["dts"], sol["ratios"]
    for k in range(len(dts) - 1):
        r = ratios[k]
        factor = 10.0 if r == 0 else min(10.0, max(0.9 / r ** 0.2, 1.0 if r < 1 else 0.2))
        nxt = dts[k] * factor
        assert dts[k + 1] <= nxt * (1 + 1e-12)            # equal unless the last step was not needed in full
    assert sum(r <= 1.0 for r in ratios) == sol["n_accepted"]


def test_dopri5_lands_on_jumps_and_reevaluates():
    calls = []

    def f(t, y, perturb=NONE):
        calls.append((float(t), perturb))
        return -y * (1.0 if float(t) < 1.0 or (float(t) == 1.0 and perturb != NEXT) else 3.0)

    y0 = torch.ones(1, 1, dtype=torch.float64)
    sol = odeint_dopri5(f, y0, [0.0, 2.0], rtol=1e-7, atol=1e-9, jump_t=[0.0, 1.0, 2.0, 3.0])
    exact = math.exp(-1.0) * math.exp(-3.0)
    assert abs(sol["ys"][-1].item() - exact) < 1e-6
    assert (1.0, NEXT) in calls                            # vector field re-evaluated just after the knot
    assert any(p == PREV for _, p in calls)                # c = 1 stages look just before the step end


@pytest.mark.parametrize("step", [None, 0.05])
def test_rk4_38_order_and_output_interpolation(step):
    A, f = _lin_system()
    y0 = torch.tensor([[1.0, 0.5]], dtype=torch.float64)
    exact = lambda T: torch.matrix_exp(A * T) @ y0[0]
    errs = []
    for n in (8, 16, 32):
        t = [2.0 * k / n for k in range(n + 1)]
        ys = odeint_rk4(f, y0, t, step_size=None)["ys"]
        errs.append((ys[-1, 0] - exact(2.0)).abs().max().item())
    assert 3.7 < math.log2(errs[0] / errs[1]) < 4.3 and 3.7 < math.log2(errs[1] / errs[2]) < 4.3
    if step:
        t = [0.0, 0.33, 1.0, 1.96, 2.0]
        ys = odeint_rk4(f, y0, t, step_size=step)["ys"]
        for k, T in enumerate(t):
            assert (ys[k, 0] - exact(T)).abs().max() < 2e-3     # linear interpolation between grid states


def test_pose_cde_reference_quirks():
    """z0 depends on the first observation; the returned state is z0 (PoseCDE.py:96,103); eval mode
    keeps a growing history and uses absolute timestamps (PoseCDE.py:81,88-92)."""
    from oracle.modules import deepvio_initialization
    from oracle.pose_cde import OraclePoseCDE
    from oracle.pose_odernn import default_opt
    torch.manual_seed(0)
    m = OraclePoseCDE(default_opt(v_f_len=8, i_f_len=8, cde_hidden_dim=16, cde_fn_num_layers=1))
    deepvio_initialization(m)
    fv, fi = 0.2 * torch.randn(3, 4, 8), 0.2 * torch.randn(3, 4, 8)
    ts = torch.arange(5.0).repeat(3, 1) * 0.1 + 7.0
    m.train()
    with torch.no_grad():
        pose, z0 = m(fv, fi, ts)
        x0 = torch.cat([(ts[:, 1:2] - ts[:, :1]), fv[:, 0], fi[:, 0]], -1)
        assert torch.allclose(z0, m.initial(x0), atol=1e-6)
        assert torch.allclose(pose[:, 0], m.regressor(z0), atol=1e-6)       # first output is z0 itself
        assert m.history is None
        m.eval()
        m(fv, fi, ts)
        assert m.history.shape == (3, 4, 17) and torch.equal(m.history[:, :, 0], ts[:, 1:])
        m(fv, fi, ts + 0.4, prev=z0)
        assert m.history.shape == (3, 8, 17)


def test_bounded_history_leaves_the_cubic_path_unchanged():
    """odevio_b200.PoseCDE's `cde_history_limit` restated in the oracle: in cubic mode a window's solve reads the control path
    on its own knots only (Hermite cubics with backward differences: one observation before the window), so truncating
    the eval-mode history (reference PoseCDE.py:88-92 grows it without bound) must not change the poses."""
    import copy
    from oracle.modules import deepvio_initialization
    from oracle.pose_cde import OraclePoseCDE
    from oracle.pose_odernn import default_opt
    opt = default_opt(v_f_len=8, i_f_len=8, cde_hidden_dim=16, cde_fn_num_layers=2, cde_interp="cubic")
    torch.manual_seed(0)
    full = OraclePoseCDE(opt)
    deepvio_initialization(full)
    full.eval()
    lim_opt = copy.copy(opt)
    lim_opt.cde_history_limit = 1                      # floor: window length + 1
    lim = OraclePoseCDE(lim_opt)
    lim.load_state_dict(full.state_dict())
    lim.eval()
    # In exact arithmetic the two are the same path; in floating point the knot index t loses ulp(t) of the in-segment
    # parameter s = t - floor(t), so the adaptive steps differ at rounding level and the solutions within the solver
    # tolerance (rtol 1e-4).  fp64: identical to 1e-9; fp32: within 2e-4 -- and the UNBOUNDED history is the less
    # accurate of the two as the knot index grows (ulp(1000) = 6e-5 in fp32).
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 2e-4)):
        a, b = copy.deepcopy(full).to(dtype), copy.deepcopy(lim).to(dtype)
        g = torch.Generator().manual_seed(1)
        B, S = 3, 4
        t0 = torch.zeros(B, 1, dtype=dtype)
        hc_a = hc_b = None
        for w in range(4):
            fv = (0.2 * torch.randn(B, S, 8, generator=g)).to(dtype)
            fi = (0.2 * torch.randn(B, S, 8, generator=g)).to(dtype)
            ts = torch.cat([t0, t0 + torch.cumsum(0.1 + 0.1 * torch.rand(B, S, generator=g), 1).to(dtype)], 1)
            t0 = ts[:, -1:]
            with torch.no_grad():
                pa, hc_a = a(fv, fi, ts, prev=hc_a)
                pb, hc_b = b(fv, fi, ts, prev=hc_b)
            assert a.history.shape[1] == (w + 1) * S and b.history.shape[1] == min((w + 1) * S, S + 1)
            assert ((pa - pb).abs().max() / pa.abs().max()).item() <= tol, (dtype, w)
