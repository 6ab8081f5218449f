"""Oracle restatement of ``PoseCDE`` (reference src/models/PoseCDE.py:41-112).  Test
infrastructure; parity unpinned (oracle/__init__.py).

Same constructor namespace, attribute names and state_dict keys as the reference
(``fuse.net.*``, ``reduction_net.{0,2}.*`` (constructed, never used: PoseCDE.py:53-57),
``initial.0.*``, ``cde_func.net.*``, ``regressor.{0,2}.*``); ``torchcde.cdeint`` is replaced by
:mod:`oracle.torchcde_like` + :mod:`oracle.torchdiffeq_like`.

Modes (optional ``opt`` attributes, reference values are the defaults):
  cde_interp = "linear"  reference: rectilinear linear path on the integer knot grid, integrated
                         over ``ts[0, 1:]`` in SECONDS (batch row 0's times; PoseCDE.py:94-101)
             = "cubic"   north_star: Hermite cubic (backward differences) through the observations
                         on knots 0..S-1, integrated over the knot grid (outputs at every knot)
  cde_atol = 1e-6, cde_rtol = 1e-4 (PoseCDE.py:101), cde_step_size = None (fixed-grid rk4)
"""

import torch
import torch.nn as nn

from .modules import OracleCDEFunc, OracleFusion, make_regressor
from .torchcde_like import HermiteCubicBackward, LinearInterpolation, linear_interpolation_coeffs
from .torchdiffeq_like import NEXT, PREV, odeint_dopri5, odeint_rk4


def _nextafter32(t, direction):
    if direction == 0:
        return t
    return torch.nextafter(t, t + direction)


class OraclePoseCDE(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.adjoint = getattr(opt, "adjoint", False)
        self.f_len = opt.v_f_len + opt.i_f_len
        self.input_dim = opt.cde_hidden_dim + 1
        self.cde_hidden_dim = opt.cde_hidden_dim
        self.fuse = OracleFusion(self.f_len, opt.fuse_method)
        self.reduction_net = nn.Sequential(nn.Linear(self.f_len, self.f_len // 2), nn.LeakyReLU(0.1, inplace=True),
                                           nn.Linear(self.f_len // 2, opt.cde_hidden_dim))
        self.initial = nn.Sequential(nn.Linear(opt.cde_hidden_dim + 1, opt.cde_hidden_dim), nn.Tanh())
        self.cde_func = OracleCDEFunc(self.input_dim, opt.cde_hidden_dim, opt.cde_fn_num_layers,
                                      opt.cde_activation_fn)
        self.regressor = make_regressor(self.cde_hidden_dim)
        self.solver = opt.cde_solver
        if self.solver not in ("dopri5", "rk4"):
            raise ValueError(f"Solver {self.solver} not supported")
        self.interp = getattr(opt, "cde_interp", "linear")
        self.atol = getattr(opt, "cde_atol", 1e-6)
        self.rtol = getattr(opt, "cde_rtol", 1e-4)
        self.step_size = getattr(opt, "cde_step_size", None)
        # restatement of odevio_b200.PoseCDE's bounded history (cubic mode only; None = the reference's unbounded growth)
        self.history_limit = getattr(opt, "cde_history_limit", None)
        self.history = None
        self.last_stats = None
        self.vf_noise = None          # tests: (eps, torch.Generator) -> k <- k * (1 + eps * N(0,1)), conditioning probe

    def forward(self, fv, fi, ts, prev=None, do_profile=False):
        fused = self.fuse(fv, fi)
        ts_diff = ts - ts[:, :1] if self.training else ts               # PoseCDE.py:81
        x = torch.cat([ts_diff[:, 1:].unsqueeze(-1), fused], dim=-1)    # channel 0 = time
        obs = x
        if not self.training:                                           # PoseCDE.py:88-92
            self.history = torch.cat([self.history, x], dim=1) if prev is not None else x
            if self.history_limit and self.interp == "cubic" and self.history.shape[1] > max(self.history_limit, x.shape[1] + 1):
                self.history = self.history[:, -max(self.history_limit, x.shape[1] + 1):]
            obs = self.history
        else:
            self.history = None
        if self.interp == "linear":
            X = LinearInterpolation(linear_interpolation_coeffs(obs, rectilinear=0))
            t_out = [float(v) for v in ts_diff[0, 1:].double()]         # batch row 0's times, seconds
        else:
            X = HermiteCubicBackward(obs)
            S = x.shape[1]
            n = obs.shape[1]
            t_out = [float(k) for k in range(n - S, n)]                 # the new observations' knots
        z0 = self.initial(X.evaluate(X.interval[0])) if prev is None else prev
        func = self.cde_func

        def vf(t, z, perturb):
            tt = _nextafter32(t.to(z.dtype), perturb)
            g = func(tt, z)                                             # [B, Hc, C]
            dX = X.derivative(tt)                                       # [B, C]
            out = (g @ dX.unsqueeze(-1)).squeeze(-1)
            if self.vf_noise is not None:
                eps, gen = self.vf_noise
                out = out * (1 + eps * torch.randn(out.shape, generator=gen, dtype=out.dtype))
            return out

        if self.solver == "dopri5":
            sol = odeint_dopri5(vf, z0, t_out, self.rtol, self.atol, jump_t=[float(v) for v in X.grid_points])
        else:
            step = self.step_size
            if step is None and self.interp == "linear":
                step = 1.0                      # torchcde injects min(diff(grid_points)) for fixed solvers
            sol = odeint_rk4(vf, z0, t_out, step_size=step)
        h = sol["ys"].transpose(0, 1)                                   # [B, S, Hc]
        self.last_stats = {k: v for k, v in sol.items() if k != "ys"}
        return self.regressor(h), z0                                    # returns z0 (PoseCDE.py:103)

    def get_reduction_net_params(self):
        return self.reduction_net.parameters()

    def get_regressor_params(self):
        return self.regressor.parameters()

    def get_other_params(self):
        return [p for n, p in self.named_parameters() if not n.startswith("regressor")]
