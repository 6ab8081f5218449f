"""Oracle restatement of torchdiffeq 0.2.3's ``odeint`` as reached through
``torchcde.cdeint`` (reference src/models/PoseCDE.py:101; solver name from
scripts/config.py:78).  Test infrastructure; parity unpinned (oracle/__init__.py).
Semantics follow SURVEY.md A.3:

  * ``dopri5``: ONE step size for the whole batch, time-like quantities in float64, state in its
    own dtype; Hairer initial step; accept iff ratio <= 1 with ratio the RMS over ALL elements;
    dt_next = dt * min(10, max(0.9 / ratio^(1/5), dfactor)), dfactor = 1 if ratio < 1 else 0.2;
    steps are shortened to land on ``jump_t`` (the control path's knots) and the vector field is
    re-evaluated just after a jump; outputs by the quartic dense output of the bracketing step.
  * ``rk4``: the 3/8 rule on a fixed grid (output times, or ``step_size`` with linear
    interpolation of the outputs).

The vector field is called as ``f(t, y, perturb)`` with ``t`` already cast to the state dtype
and perturb in {-1, 0, +1} (torchdiffeq's Perturb.PREV / NONE / NEXT: one ulp before / after).
Elementwise arithmetic is written as explicit chains so the CUDA kernel can mirror it.
"""

from typing import Callable, Dict, List, Optional

import torch

from .tableaus import DOPRI5

PREV, NONE, NEXT = -1, 0, 1


def _rms(x):
    return torch.sqrt(torch.mean(x.double() * x.double())).item()      # python float (fp64)


def _wsum_dt(coeffs, ks, dt_s):
    """sum_j k_j * fl(c_j * dt) left to right, skipping exact zeros (torchdiffeq: k.matmul(beta * dt))."""
    acc = None
    for cj, kj in zip(coeffs, ks):
        if cj == 0.0:
            continue
        term = kj * (torch.tensor(cj, dtype=kj.dtype) * dt_s)
        acc = term if acc is None else acc + term
    return acc


def select_initial_step(f, t0, y0, order, rtol, atol, f0):
    """Hairer's rule (torchdiffeq ``_select_initial_step``); returns a python float and the
    number of extra vector-field evaluations (1 unless the degenerate branch is taken)."""
    dtype = y0.dtype
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = 1e-6
    else:
        h0 = 0.01 * d0 / d1
    h0_s = torch.tensor(h0, dtype=dtype)
    y1 = y0 + h0_s * f0
    f1 = f(torch.tensor(t0 + h0, dtype=dtype), y1, NONE)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    return min(100 * h0, h1), 1


def interp_fit(y0, y1, y_mid, f0, f1, dt_s):
    a = 2.0 * dt_s * (f1 - f0) - 8.0 * (y1 + y0) + 16.0 * y_mid
    b = dt_s * (5.0 * f0 - 3.0 * f1) + 18.0 * y0 + 14.0 * y1 - 32.0 * y_mid
    c = dt_s * (f1 - 4.0 * f0) - 11.0 * y0 - 5.0 * y1 + 16.0 * y_mid
    d = dt_s * f0
    return [y0, d, c, b, a]                       # ascending powers, evaluated as a power sum


def interp_evaluate(coeffs, t0, t1, t, dtype):
    x = torch.tensor((t - t0) / (t1 - t0), dtype=dtype)
    total = coeffs[0] + x * coeffs[1]
    xp = x
    for cf in coeffs[2:]:
        xp = xp * x
        total = total + xp * cf
    return total


def odeint_dopri5(f: Callable, y0: torch.Tensor, t: List[float], rtol: float, atol: float,
                  jump_t: Optional[List[float]] = None, max_num_steps: int = 100000,
                  first_step: Optional[float] = None) -> Dict:
    """Returns ys [len(t), *y0.shape] and stats.  ``t`` ascending python floats (fp64)."""
    tab = DOPRI5
    dtype = y0.dtype
    t = [float(v) for v in t]
    jump = sorted(float(v) for v in (jump_t or []))
    import bisect
    next_jump = min(bisect.bisect_right(jump, t[0]), len(jump) - 1) if jump else -1
    f0 = f(torch.tensor(t[0], dtype=dtype), y0, NONE)
    n_f = 1
    if first_step is None:
        dt, extra = select_initial_step(f, t[0], y0, tab.order - 1, rtol, atol, f0)
        n_f += extra
    else:
        dt = float(first_step)
    t0 = t1 = t[0]
    y = y0
    coeffs = [y0] * 5
    ys = [y0]
    n_steps = n_acc = 0
    dts, ratios = [], []
    for next_t in t[1:]:
        while next_t > t1:
            if n_steps >= max_num_steps:
                raise RuntimeError("max_num_steps exceeded")
            # ---- one adaptive step from (t1, y, f0)
            ta = t1
            tb = ta + dt
            on_jump = False
            step = dt
            if jump and next_jump >= 0:
                nj = jump[next_jump]
                on_jump = ta < nj < ta + step
                if on_jump:
                    step = nj - ta
                    tb = nj
            dt_s = torch.tensor(step, dtype=dtype)         # t0, dt, t1 cast to the state dtype
            ta_s, tb_s = torch.tensor(ta, dtype=dtype), torch.tensor(tb, dtype=dtype)
            ks = [f0]
            yi = y
            for i in range(1, tab.n_stages):
                yi = y + _wsum_dt(tab.a[i], ks, dt_s)
                if tab.c[i] == 1.0:
                    ks.append(f(tb_s, yi, PREV))
                else:
                    ks.append(f(ta_s + torch.tensor(tab.c[i], dtype=dtype) * dt_s, yi, NONE))
            n_f += tab.n_stages - 1
            y1, f1 = yi, ks[-1]                             # dopri5: y1 is the last stage's argument
            err = _wsum_dt(tab.e, ks, dt_s)
            tol = atol + rtol * torch.maximum(y.abs(), y1.abs())
            ratio = _rms(err / tol)
            accept = ratio <= 1.0
            n_steps += 1
            dts.append(step); ratios.append(ratio)
            if accept:
                n_acc += 1
                y_mid = y + _wsum_dt(tab.b_mid, ks, dt_s)
                coeffs = interp_fit(y, y1, y_mid, ks[0], f1, dt_s)
                t0, t1, y = ta, tb, y1
                if on_jump:
                    if next_jump != len(jump) - 1:
                        next_jump += 1
                    f1 = f(tb_s, y1, NEXT)                  # vector field just after the jump
                    n_f += 1
                f0 = f1
            # step-size update (torchdiffeq _optimal_step_size)
            if ratio == 0.0:
                dt = step * 10.0
            else:
                dfactor = 1.0 if ratio < 1.0 else 0.2
                factor = min(10.0, max(0.9 / ratio ** 0.2, dfactor))
                dt = step * factor
        ys.append(interp_evaluate(coeffs, t0, t1, next_t, dtype))
    return dict(ys=torch.stack(ys, 0), n_steps=n_steps, n_accepted=n_acc, n_f_evals=n_f, dts=dts, ratios=ratios)


def rk4_38_step(f, ta_s, dt_s, tb_s, y):
    k1 = f(ta_s, y, NONE)
    third = torch.tensor(1.0 / 3.0, dtype=y.dtype)
    k2 = f(ta_s + dt_s * third, y + dt_s * k1 * third, NONE)
    k3 = f(ta_s + dt_s * (2.0 * third), y + dt_s * (k2 - k1 * third), NONE)
    k4 = f(tb_s, y + dt_s * (k1 - k2 + k3), PREV)
    return y + dt_s * (k1 + 3.0 * (k2 + k3) + k4) * 0.125


def odeint_rk4(f: Callable, y0: torch.Tensor, t: List[float], step_size: Optional[float] = None) -> Dict:
    """torchdiffeq fixed-grid ``rk4`` (3/8 rule).  Without step_size the grid is the output
    times; with it, grid = t[0] + step_size * arange(...) ending at t[-1] and outputs are linear
    interpolations between grid states."""
    import math
    dtype = y0.dtype
    t = [float(v) for v in t]
    if step_size is None:
        grid = t
    else:
        niters = math.ceil((t[-1] - t[0]) / step_size + 1)
        grid = [t[0] + step_size * k for k in range(niters)]
        grid[-1] = t[-1]
    ys = [y0]
    j = 1
    y = y0
    n_f = 0
    for ta, tb in zip(grid[:-1], grid[1:]):
        dt_s = torch.tensor(tb - ta, dtype=dtype)
        y1 = rk4_38_step(f, torch.tensor(ta, dtype=dtype), dt_s, torch.tensor(tb, dtype=dtype), y)
        n_f += 4
        while j < len(t) and tb >= t[j]:
            if tb == t[j]:
                ys.append(y1)
            else:
                w = torch.tensor((t[j] - ta) / (tb - ta), dtype=dtype)
                ys.append(y + w * (y1 - y))
            j += 1
        y = y1
    return dict(ys=torch.stack(ys, 0), n_steps=len(grid) - 1, n_accepted=len(grid) - 1, n_f_evals=n_f)
