"""Oracle restatement of ``PoseODERNN`` (reference src/models/PoseODERNN.py:39-123).
Test infrastructure; parity unpinned (oracle/__init__.py).

Same constructor namespace (``opt``), attribute names and state_dict keys as the
reference (``ode_func.net.*``, ``rnn.*``, ``fuse.net.*``, ``regressor.*``); the
torchode solve is replaced by :mod:`oracle.torchode_like`.  Extra optional
``opt`` attributes (reference values are the defaults): ``ode_atol`` 1e-6,
``ode_rtol`` 1e-2, ``ode_dt0`` 1e-4 (PoseODERNN.py:57,72), ``ode_substeps`` 1.
Solver menu = the reference's {dopri5, heun, tsit5, euler} (PoseODERNN.py:125-137)
plus north_star's fixed-step {rk4, rk4_38}.
"""

from types import SimpleNamespace

import torch
import torch.nn as nn

from . import tableaus
from .modules import OracleODEFunc, OracleFusion, make_regressor
from .torchode_like import ControllerOptions, solve_adaptive, solve_fixed

ADAPTIVE = ("dopri5", "heun", "tsit5", "euler")
FIXED = ("rk4", "rk4_38")


def default_opt(**over):
    """The hot-path subset of reference scripts/config.py:29-79 defaults."""
    d = dict(v_f_len=512, i_f_len=256, fuse_method="cat", ode_hidden_dim=512,
             ode_fn_num_layers=3, ode_activation_fn="tanh", ode_solver="dopri5",
             ode_rnn_type="rnn", rnn_num_layers=2, rnn_hidden_dim=1024, rnn_dropout_out=0.0,
             cde_hidden_dim=128, cde_fn_num_layers=3, cde_num_layers=3,
             cde_activation_fn="tanh", cde_solver="dopri5", adjoint=False, seq_len=11,
             model_type="ode-rnn")
    d.update(over)
    return SimpleNamespace(**d)


class OraclePoseODERNN(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.f_len = opt.v_f_len + opt.i_f_len
        self.rnn_num_layers = opt.rnn_num_layers
        self.ode_func = OracleODEFunc(self.f_len, opt.ode_hidden_dim, opt.ode_fn_num_layers,
                                      opt.ode_activation_fn)
        if opt.ode_solver not in ADAPTIVE + FIXED:
            raise ValueError(f"Solver {opt.ode_solver} not supported")
        self.solver_name = opt.ode_solver
        self.ctrl = ControllerOptions(atol=getattr(opt, "ode_atol", 1e-6),
                                      rtol=getattr(opt, "ode_rtol", 1e-2),
                                      accept_strict=getattr(opt, "ode_accept_strict", True),
                                      floor_factor_after_accept=getattr(opt, "ode_floor_factor", False),
                                      endpoint=getattr(opt, "ode_endpoint", "y1"),
                                      exact_landing=getattr(opt, "ode_exact_landing", True),
                                      max_steps=getattr(opt, "ode_max_steps", 100000),
                                      detach_dt=getattr(opt, "ode_detach_dt", False))
        self.dt0 = getattr(opt, "ode_dt0", 1e-4)
        self.substeps = getattr(opt, "ode_substeps", 1)
        self.trace_steps = getattr(opt, "ode_trace_steps", 0)
        if opt.ode_rnn_type == "rnn":
            self.rnn = nn.RNN(self.f_len, self.f_len, num_layers=opt.rnn_num_layers, batch_first=True)
        elif opt.ode_rnn_type == "gru":
            self.rnn = nn.GRU(self.f_len, self.f_len, num_layers=opt.rnn_num_layers, batch_first=True)
        else:
            raise ValueError(f"RNN type {opt.ode_rnn_type} not supported")
        self.fuse = OracleFusion(self.f_len, opt.fuse_method)
        self.regressor = make_regressor(self.f_len)
        self.last_stats = None

    def evolve_state(self, state, ts2):
        """PoseODERNN.evolve_state (PoseODERNN.py:70-75) for one layer's hidden state."""
        f = self.ode_func
        if self.solver_name in FIXED:
            return solve_fixed(f, state, ts2, tableaus.BY_NAME[self.solver_name], self.substeps)
        dt0 = torch.full((ts2.shape[0],), self.dt0, dtype=ts2.dtype, device=ts2.device)
        return solve_adaptive(f, state, ts2, dt0, tableaus.BY_NAME[self.solver_name], self.ctrl)

    def forward(self, fv, fi, ts, prev=None, do_profile=False):
        fused = self.fuse(fv, fi)
        B, S, _ = fused.shape
        h = (torch.zeros(self.rnn_num_layers, B, self.f_len, dtype=fused.dtype, device=fused.device)
             if prev is None else prev)
        ts_diff = ts - ts[:, :1] if prev is None else ts          # PoseODERNN.py:100
        outs = []
        L = self.rnn_num_layers
        n_steps = torch.zeros(S, L, B, dtype=torch.int64)
        n_acc = torch.zeros(S, L, B, dtype=torch.int64)
        n_f = 0
        T = self.trace_steps
        trace = torch.zeros(S, L, B, T, 2, dtype=torch.float32) if T else None
        for i in range(S):
            evolved = []
            for j in range(L):
                sol = self.evolve_state(h[j], ts_diff[:, i:i + 2])
                evolved.append(sol["y_end"])
                n_steps[i, j], n_acc[i, j] = sol["n_steps"], sol["n_accepted"]
                n_f += sol["n_f_evals"]
                if T and "trace_dt" in sol:
                    running = torch.ones(B, dtype=torch.bool)
                    for k in range(min(T, len(sol["trace_dt"]))):   # only steps the row actually took
                        took = sol["n_steps"] > k
                        trace[i, j, :, k, 0] = torch.where(took, sol["trace_dt"][k].float(), trace[i, j, :, k, 0])
                        trace[i, j, :, k, 1] = torch.where(took, sol["trace_ratio"][k].float(), trace[i, j, :, k, 1])
            out_i, h = self.rnn(fused[:, i:i + 1, :], torch.stack(evolved, 0))
            outs.append(out_i)
        pose = self.regressor(torch.cat(outs, 1))
        self.last_stats = dict(n_steps=n_steps, n_accepted=n_acc, n_f_evals=n_f, trace=trace)
        return pose, h

    def get_regressor_params(self):
        return self.regressor.parameters()

    def get_other_params(self):
        return [p for n, p in self.named_parameters() if not n.startswith("regressor")]
