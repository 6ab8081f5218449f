"""Oracle restatement of torchode 0.2.0's batched explicit-RK solve loop with an
*independent step size, accept decision and status per batch row* (test
infrastructure; parity unpinned -- see oracle/__init__.py).

Call site being restated: ``PoseODERNN.evolve_state``
(reference src/models/PoseODERNN.py:70-75) =
``to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(atol=1e-6, rtol=1e-2, term))
.solve(to.InitialValueProblem(y0=state, t_eval=ts[:, i:i+2]), dt0=1e-4).ys[:, -1]``
(construction at src/models/PoseODERNN.py:55-60).  Semantics follow SURVEY.md A.1.

The elementwise arithmetic is written as explicit two-rounding mul/add chains
(no addcmul/einsum) so the CUDA kernels can mirror the operation order.
"""

from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

from .tableaus import Tableau, DOPRI5

STATUS_OK, STATUS_MAX_STEPS, STATUS_INFINITE_NORM = 0, 1, 2


@dataclass
class ControllerOptions:
    """torchode IntegralController == PIDController(pcoeff=0, icoeff=1, dcoeff=0)."""
    atol: float = 1e-6                    # PoseODERNN.py:57
    rtol: float = 1e-2                    # PoseODERNN.py:57
    safety: float = 0.9
    factor_min: float = 0.2
    factor_max: float = 10.0
    accept_strict: bool = True            # accept iff ratio < 1 (A.1, confidence M)
    floor_factor_after_accept: bool = False   # torchdiffeq/diffrax rule; torchode believed not (L)
    # End point of the interval.  torchode (as recalled, confidence M) evaluates the step's
    # quartic dense output at t_end.  In this call pattern (t_eval = interval ends only, dt clamped
    # to land on t_end) that is ALWAYS x = (t_end - t0)/dt = 1, where the quartic equals y1 in exact
    # arithmetic but carries ~32|y| ulp of cancellation noise in fp32 (hard part 4).
    # "y1" returns that exact-arithmetic value; "dense" reproduces the literal fp32 Horner form.
    endpoint: str = "y1"
    # When the controller clamps dt to the remaining interval the step ends at t_end.  Literal fp32
    # arithmetic computes fl(t + fl(t_end - t)), which can fall one ulp short and trigger an extra
    # one-ulp step whose occurrence depends on rounding noise; exact_landing sets t = t_end.
    exact_landing: bool = True
    max_steps: int = 100000               # torchode max_steps=None; guard so nothing spins forever
    detach_dt: bool = False               # True: treat step sizes as constants under autograd
                                          # (torchode backprop_through_step_size_control=True => False here)


def _weighted_sum(coeffs, ks, dtype):
    """acc = c0*k0 + c1*k1 + ... left to right, skipping exact-zero coefficients."""
    acc = None
    for cj, kj in zip(coeffs, ks):
        if cj == 0.0:
            continue
        term = kj * torch.tensor(cj, dtype=dtype)
        acc = term if acc is None else acc + term
    return acc


def rk_step(f: Callable, tab: Tableau, t0, y0, dt, k0=None):
    """One explicit RK step for every row.  Returns (y1, err | None, ks)."""
    dtype = y0.dtype
    dtc = dt.to(dtype)[:, None]
    ks = [f(t0, y0) if k0 is None else k0]
    y_i = y0
    for i in range(1, tab.n_stages):
        y_i = y0 + dtc * _weighted_sum(tab.a[i], ks, dtype)
        ks.append(f(t0 + tab.c[i] * dt, y_i))
    if tab.ssal:
        y1 = y_i
    else:
        y1 = y0 + dtc * _weighted_sum(tab.b, ks, dtype)
    err = None
    if tab.e is not None:
        err = dtc * _weighted_sum(tab.e, ks, dtype)
    return y1, err, ks


def error_ratio(err, y0, y1, atol, rtol):
    """Per-row RMS of err / (atol + rtol * max(|y0|, |y1|))."""
    bound = atol + rtol * torch.maximum(y0.abs(), y1.abs())
    r = err / bound
    return torch.sqrt(torch.mean(r * r, dim=-1))


def dense_eval(tab: Tableau, x, y0, y1, dt, ks):
    """Value of the step's interpolant at x = (t - t0)/dt."""
    dtype = y0.dtype
    x = x.to(dtype)[:, None]
    if tab.b_mid is None:                       # heun / euler: linear
        return y0 + x * (y1 - y0)
    dtc = dt.to(dtype)[:, None]
    f0 = dtc * ks[0]
    f1 = dtc * ks[-1]
    y_mid = y0 + dtc * _weighted_sum(tab.b_mid, ks, dtype)
    a = 2.0 * (f1 - f0) - 8.0 * (y1 + y0) + 16.0 * y_mid
    b = 5.0 * f0 - 3.0 * f1 + 18.0 * y0 + 14.0 * y1 - 32.0 * y_mid
    c = f1 - 4.0 * f0 - 11.0 * y0 - 5.0 * y1 + 16.0 * y_mid
    return (((a * x + b) * x + c) * x + f0) * x + y0


def solve_adaptive(f: Callable, y0: torch.Tensor, t_eval: torch.Tensor, dt0: torch.Tensor,
                   tab: Tableau = DOPRI5, opts: Optional[ControllerOptions] = None) -> Dict:
    """Integrate every row b from t_eval[b,0] to t_eval[b,1]; returns the state at
    t_eval[b,1] (``ys[:, -1, :]`` of the reference) and per-row statistics."""
    opts = opts or ControllerOptions()
    B = y0.shape[0]
    tdt = t_eval.dtype
    t_start, t_end = t_eval[:, 0].clone(), t_eval[:, 1].clone()
    t_min, t_max = torch.minimum(t_start, t_end), torch.maximum(t_start, t_end)
    t, y = t_start.clone(), y0.clone()
    y_end = y0.clone()                                   # rows with t_start == t_end keep y0
    n_steps = torch.zeros(B, dtype=torch.int64)
    n_acc = torch.zeros(B, dtype=torch.int64)
    status = torch.zeros(B, dtype=torch.int64)
    n_f = 0

    k0 = None
    if tab.fsal:
        k0 = f(t, y)
        n_f += 1
    dt = torch.clamp(dt0.to(tdt).clone(), t_min - t, t_max - t)
    running = t < t_end
    not_evaluated = running.clone()
    exponent = -1.0 / tab.order
    loops = 0
    trace_dt, trace_ratio = [], []
    while bool(running.any()):
        loops += 1
        y1, err, ks = rk_step(f, tab, t, y, dt, k0)
        n_f += tab.n_stages - (1 if tab.fsal else 0)
        trace_dt.append(dt.detach().clone())
        if err is None:
            trace_ratio.append(torch.zeros(B, dtype=tdt))
            accept = torch.ones(B, dtype=torch.bool)
            dt_next = dt.clone()
            finite = torch.ones(B, dtype=torch.bool)
        else:
            ratio = error_ratio(err, y, y1, opts.atol, opts.rtol).to(tdt)
            finite = torch.isfinite(ratio)
            trace_ratio.append(ratio.detach().clone())
            accept = (ratio < 1.0) if opts.accept_strict else (ratio <= 1.0)
            factor = opts.safety * torch.pow(ratio, exponent)
            factor = torch.clamp(factor, opts.factor_min, opts.factor_max)
            if opts.floor_factor_after_accept:
                factor = torch.where(accept, torch.clamp(factor, min=1.0), factor)
            dt_next = dt * factor
            if opts.detach_dt:
                dt_next = dt_next.detach()
        upd = accept & running
        n_steps += running
        n_acc += upd
        t_next = t + dt
        if opts.exact_landing:
            t_next = torch.where(dt >= t_end - t, t_end, t_next)
        t_new = torch.where(upd, t_next, t)
        # dense output at t_end for rows that just reached / passed it
        to_eval = upd & (t_new >= t_end) & not_evaluated
        if bool(to_eval.any()):
            if opts.endpoint == "dense":
                x = (t_end - t) / dt
                val = dense_eval(tab, x, y, y1, dt, ks)
            else:
                val = y1
            y_end = torch.where(to_eval[:, None], val, y_end)
            not_evaluated = not_evaluated & ~to_eval
        y = torch.where(upd[:, None], y1, y)
        if tab.fsal:
            k0 = torch.where(upd[:, None], ks[-1], k0)
        t = t_new
        bad = running & ~finite
        status = torch.where(bad, torch.full_like(status, STATUS_INFINITE_NORM), status)
        running = running & (t < t_end) & finite
        if loops >= opts.max_steps:
            status = torch.where(running, torch.full_like(status, STATUS_MAX_STEPS), status)
            break
        dt = torch.where(running, dt_next, dt)
        dt = torch.clamp(dt, t_min - t, t_max - t)
    # rows that failed (max_steps / non-finite norm) return their last accepted state
    y_end = torch.where((status != STATUS_OK)[:, None], y, y_end)
    return dict(y_end=y_end, n_steps=n_steps, n_accepted=n_acc, status=status,
                n_f_evals=n_f, loops=loops, trace_dt=trace_dt, trace_ratio=trace_ratio)


def solve_fixed(f: Callable, y0: torch.Tensor, t_eval: torch.Tensor, tab: Tableau,
                substeps: int = 1) -> Dict:
    """Fixed-step integration with ``substeps`` equal steps per row over
    [t_eval[b,0], t_eval[b,1]] (north_star "fixed-step rk4"; not in the
    reference's ODE-RNN menu, src/models/PoseODERNN.py:125-137)."""
    tdt = t_eval.dtype
    t0 = t_eval[:, 0].clone()
    h = (t_eval[:, 1] - t_eval[:, 0]) / torch.tensor(float(substeps), dtype=tdt)
    y = y0.clone()
    for s in range(substeps):
        ts = t0 + h * float(s)
        y, _, _ = rk_step(f, tab, ts, y, h, None)
    B = y0.shape[0]
    steps = torch.full((B,), substeps, dtype=torch.int64)
    return dict(y_end=y, n_steps=steps, n_accepted=steps.clone(),
                status=torch.zeros(B, dtype=torch.int64), n_f_evals=substeps * tab.n_stages,
                loops=substeps)
