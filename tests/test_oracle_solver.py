"""The oracle's solver restatement against independent anchors (SciPy RK45 / analytic solutions)
and its own invariants.  CPU only."""

import copy

import numpy as np
import pytest
import torch
from scipy.integrate import solve_ivp

from oracle import tableaus as T
from oracle.torchode_like import ControllerOptions, dense_eval, rk_step, solve_adaptive, solve_fixed


def _linear_field(D, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    A = scale * torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5
    return A, (lambda t, y: y @ A.to(y.dtype).T)


def test_dopri5_converges_to_scipy_solution():
    D, B = 6, 5
    A, f = _linear_field(D, scale=2.0)
    g = torch.Generator().manual_seed(1)
    y0 = torch.randn(B, D, generator=g, dtype=torch.float64)
    t_eval = torch.tensor([[0.0, 0.1 * (b + 1)] for b in range(B)], dtype=torch.float64)
    opts = ControllerOptions(atol=1e-10, rtol=1e-9)
    sol = solve_adaptive(f, y0, t_eval, torch.full((B,), 1e-4, dtype=torch.float64), T.DOPRI5, opts)
    for b in range(B):
        ref = solve_ivp(lambda t, y: A.numpy() @ y, (0.0, t_eval[b, 1].item()), y0[b].numpy(), method="RK45",
                        rtol=1e-12, atol=1e-13).y[:, -1]
        assert np.allclose(sol["y_end"][b].numpy(), ref, rtol=1e-7, atol=1e-9)
    assert (sol["status"] == 0).all()


@pytest.mark.parametrize("name,order", [("rk4", 4), ("rk4_38", 4), ("euler", 1), ("heun", 2), ("dopri5", 5), ("tsit5", 5)])
def test_single_step_order_of_accuracy(name, order):
    """Local error of one step scales like h^(order+1) on y' = -y."""
    tab = T.BY_NAME[name]
    f = lambda t, y: -y
    y0 = torch.ones(1, 1, dtype=torch.float64)
    errs = []
    for h in (0.2, 0.1):
        y1, _, _ = rk_step(f, tab, torch.zeros(1, dtype=torch.float64), y0, torch.full((1,), h, dtype=torch.float64))
        errs.append(abs(y1.item() - np.exp(-h)))
    assert abs(np.log2(errs[0] / errs[1]) - (order + 1)) < 0.5   # tsit5's tiny error constant -> slower asymptote


@pytest.mark.parametrize("name", ["dopri5", "tsit5"])
def test_dense_output_is_fourth_order_and_hits_endpoints(name):
    tab = T.BY_NAME[name]
    A, f = _linear_field(4, seed=3)
    y0 = torch.randn(2, 4, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    t0 = torch.zeros(2, dtype=torch.float64)
    dt = torch.full((2,), 0.05, dtype=torch.float64)
    y1, _, ks = rk_step(f, tab, t0, y0, dt)
    assert torch.allclose(dense_eval(tab, torch.zeros(2, dtype=torch.float64), y0, y1, dt, ks), y0, atol=1e-14)
    assert torch.allclose(dense_eval(tab, torch.ones(2, dtype=torch.float64), y0, y1, dt, ks), y1, atol=1e-13)
    mid = dense_eval(tab, torch.full((2,), 0.5, dtype=torch.float64), y0, y1, dt, ks)
    exact = y0 @ torch.linalg.matrix_exp(A * 0.025).T
    assert (mid - exact).abs().max() < 1e-8


def test_per_row_independence_and_masking():
    """Rows are integrated independently: solving a sub-batch gives identical rows, finished rows
    are not disturbed by rows that keep stepping."""
    D, B = 8, 6
    _, f = _linear_field(D, seed=5, scale=3.0)
    y0 = torch.randn(B, D, generator=torch.Generator().manual_seed(4))
    t_eval = torch.tensor([[0.0, 0.05 * (1 + 3 * b)] for b in range(B)])
    dt0 = torch.full((B,), 1e-4)
    full = solve_adaptive(f, y0, t_eval, dt0, T.DOPRI5, ControllerOptions(rtol=1e-4))
    part = solve_adaptive(f, y0[:2], t_eval[:2], dt0[:2], T.DOPRI5, ControllerOptions(rtol=1e-4))
    assert torch.equal(full["y_end"][:2], part["y_end"])
    assert torch.equal(full["n_steps"][:2], part["n_steps"])
    assert (full["n_steps"][1:] >= full["n_steps"][:-1]).all()


def test_zero_length_interval_and_landing():
    _, f = _linear_field(4, seed=6)
    y0 = torch.randn(3, 4, generator=torch.Generator().manual_seed(8))
    t_eval = torch.tensor([[0.3, 0.3], [0.0, 0.1], [100.0, 100.1]])
    sol = solve_adaptive(f, y0, t_eval, torch.full((3,), 1e-4), T.DOPRI5, ControllerOptions())
    assert torch.equal(sol["y_end"][0], y0[0]) and sol["n_steps"][0] == 0
    assert (sol["n_steps"][1:] > 0).all() and (sol["status"] == 0).all()
    # literal fp32 landing can need one extra 1-ulp step; exact landing never does
    lit = solve_adaptive(f, y0, t_eval, torch.full((3,), 1e-4), T.DOPRI5, ControllerOptions(exact_landing=False))
    assert (lit["n_steps"] >= sol["n_steps"]).all() and (lit["n_steps"] - sol["n_steps"]).max() <= 1


def test_endpoint_modes_agree_to_rounding():
    _, f = _linear_field(16, seed=9)
    y0 = torch.randn(4, 16, generator=torch.Generator().manual_seed(10))
    t_eval = torch.tensor([[0.0, 0.1]] * 4)
    a = solve_adaptive(f, y0, t_eval, torch.full((4,), 1e-4), T.DOPRI5, ControllerOptions(endpoint="y1"))
    b = solve_adaptive(f, y0, t_eval, torch.full((4,), 1e-4), T.DOPRI5, ControllerOptions(endpoint="dense"))
    assert torch.equal(a["n_steps"], b["n_steps"])
    assert (a["y_end"] - b["y_end"]).abs().max() < 64 * 2 ** -23 * y0.abs().max()   # ~32|y| ulp of cancellation


def test_failure_statuses_do_not_hang():
    blow = lambda t, y: y * y * 1e6
    y0 = torch.ones(2, 3)
    t_eval = torch.tensor([[0.0, 1.0]] * 2)
    sol = solve_adaptive(blow, y0, t_eval, torch.full((2,), 1e-4), T.DOPRI5, ControllerOptions(max_steps=50))
    assert (sol["status"] != 0).all() and sol["loops"] <= 50


def test_fixed_step_matches_manual_rk4():
    _, f = _linear_field(5, seed=11)
    y0 = torch.randn(3, 5, generator=torch.Generator().manual_seed(12), dtype=torch.float64)
    t_eval = torch.tensor([[0.0, 0.2], [0.1, 0.4], [1.0, 1.05]], dtype=torch.float64)
    sol = solve_fixed(f, y0, t_eval, T.RK4, substeps=2)
    h = (t_eval[:, 1] - t_eval[:, 0])[:, None] / 2
    y = y0.clone()
    for _ in range(2):
        k1 = f(0, y); k2 = f(0, y + h * k1 / 2); k3 = f(0, y + h * k2 / 2); k4 = f(0, y + h * k3)
        y = y + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6
    assert torch.allclose(sol["y_end"], y, rtol=1e-13, atol=1e-14)


def test_autograd_through_solver_matches_finite_differences():
    torch.manual_seed(0)
    W = torch.nn.Parameter(0.3 * torch.randn(4, 4, dtype=torch.float64))
    f = lambda t, y: torch.tanh(y @ W.T)
    y0 = torch.randn(2, 4, dtype=torch.float64)
    t_eval = torch.tensor([[0.0, 0.2]] * 2, dtype=torch.float64)
    opts = ControllerOptions(rtol=1e-6, atol=1e-8, detach_dt=True)
    loss = solve_adaptive(f, y0, t_eval, torch.full((2,), 1e-3, dtype=torch.float64), T.DOPRI5, opts)["y_end"].sum()
    (g,) = torch.autograd.grad(loss, W)
    eps = 1e-6
    Wp = W.detach().clone(); Wp[1, 2] += eps
    fp = lambda t, y: torch.tanh(y @ Wp.T)
    lp = solve_adaptive(fp, y0, t_eval, torch.full((2,), 1e-3, dtype=torch.float64), T.DOPRI5, opts)["y_end"].sum()
    assert abs((lp - loss).item() / eps - g[1, 2].item()) < 1e-4
