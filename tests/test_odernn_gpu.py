"""GPU parity: fused sm_100a ODE-RNN forward (through the C ABI) vs the CPU oracle on the same
seeded weights / inputs / timestamps.  Tolerance: max-norm relative pose error <= 1e-5 (fp32),
identical per-row step counts (n_steps, n_accepted) for the adaptive solvers."""

import pytest
import torch

from helpers import POSE_RTOL, STATE_RTOL, inputs, make_pair, run_pair

pytestmark = pytest.mark.gpu


def _check(out, steps=True):
    assert out["status_max"] == 0
    assert out["pose_err"] <= POSE_RTOL, out
    assert out["h_err"] <= STATE_RTOL, out
    if steps:
        assert out["steps_equal"] and out["acc_equal"], out


def test_config1_rk4(cuda_device):
    """BASELINE config 1: fixed-step rk4, B=16, seq_len 11."""
    ref, mod = make_pair(cuda_device, ode_solver="rk4")
    out = run_pair(ref, mod, *inputs(16))
    _check(out)


@pytest.mark.parametrize("solver", ["rk4_38", "dopri5", "tsit5", "heun"])
def test_solver_menu(cuda_device, solver):
    ref, mod = make_pair(cuda_device, ode_solver=solver, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(16, irregular=True))
    _check(out)


def test_euler_fixed_dt(cuda_device):
    # torchode Euler keeps dt = dt0 -> 1e-4 steps would be 1000 steps/interval; use a coarser dt0
    ref, mod = make_pair(cuda_device, ode_solver="euler", ode_dt0=0.02, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(8, S=4))
    _check(out)


def test_dopri5_irregular_rtol(cuda_device):
    """BASELINE config 2 semantics at a size the oracle finishes quickly."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(64, irregular=True, seed=3))
    _check(out)


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "softplus"])
def test_activations(cuda_device, act):
    ref, mod = make_pair(cuda_device, ode_activation_fn=act, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(8, S=5, irregular=True))
    _check(out)


def test_gru_jump(cuda_device):
    ref, mod = make_pair(cuda_device, ode_rnn_type="gru", bias_std=0.05)
    out = run_pair(ref, mod, *inputs(16, irregular=True))
    _check(out)


@pytest.mark.parametrize("L,H,n", [(1, 128, 2), (3, 1024, 2), (4, 256, 1), (2, 512, 4)])
def test_shapes(cuda_device, L, H, n):
    ref, mod = make_pair(cuda_device, rnn_num_layers=L, ode_hidden_dim=H, ode_fn_num_layers=n, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(11, S=4, irregular=True))      # B not a multiple of the tile
    _check(out)


def test_rows16_tile(cuda_device):
    ref, mod = make_pair(cuda_device, ode_rows_per_tile=16, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(24, S=4, irregular=True))
    _check(out)


def test_prev_carry_absolute_time(cuda_device):
    """hc carried across windows with ABSOLUTE timestamps (KITTI_eval.py:141, PoseODERNN.py:97-100)."""
    ref, mod = make_pair(cuda_device, bias_std=0.05)
    fv, fi, ts = inputs(8, S=6, irregular=True, offset=123.0)
    g = torch.Generator().manual_seed(5)
    prev = 0.5 * torch.randn(2, 8, 768, generator=g)
    out = run_pair(ref, mod, fv, fi, ts, prev=prev)
    _check(out)


def test_two_windows_equal_one(cuda_device):
    """Carrying h across two S=5 windows == one S=10 window (absolute time, same grid)."""
    ref, mod = make_pair(cuda_device, bias_std=0.05)
    fv, fi, ts = inputs(8, S=10, irregular=True, offset=50.0)
    dev = cuda_device
    h0 = torch.zeros(2, 8, 768)
    with torch.no_grad():
        p_all, h_all = mod(fv.to(dev), fi.to(dev), ts.to(dev), prev=h0.to(dev))
        p_a, h_a = mod(fv[:, :5].to(dev), fi[:, :5].to(dev), ts[:, :6].to(dev), prev=h0.to(dev))
        p_b, h_b = mod(fv[:, 5:].to(dev), fi[:, 5:].to(dev), ts[:, 5:].to(dev), prev=h_a)
    assert torch.equal(torch.cat([p_a, p_b], 1), p_all)
    assert torch.equal(h_b, h_all)


def test_shard_invariance(cuda_device):
    """Rows are independent: a batch split in two gives bit-identical rows."""
    ref, mod = make_pair(cuda_device, bias_std=0.05)
    fv, fi, ts = inputs(40, S=4, irregular=True)
    dev = cuda_device
    with torch.no_grad():
        p, h = mod(fv.to(dev), fi.to(dev), ts.to(dev))
        p1, h1 = mod(fv[:24].to(dev), fi[:24].to(dev), ts[:24].to(dev))
        p2, h2 = mod(fv[24:].to(dev), fi[24:].to(dev), ts[24:].to(dev))
    assert torch.equal(torch.cat([p1, p2], 0), p)
    assert torch.equal(torch.cat([h1, h2], 1), h)


def test_zero_length_intervals_is_plain_rnn(cuda_device):
    """All timestamps equal -> no ODE evolution -> PoseRNN (reference src/models/PoseRNN.py:64-68)."""
    ref, mod = make_pair(cuda_device, bias_std=0.05)
    fv, fi, ts = inputs(8, S=5)
    ts = torch.zeros_like(ts)
    with torch.no_grad():
        fused = torch.cat([fv, fi], -1)
        out_r, h_r = ref.rnn(fused)
        pose_r = ref.regressor(out_r)
        p, h = mod(fv.to(cuda_device), fi.to(cuda_device), ts.to(cuda_device))
    assert ((p.cpu() - pose_r).abs().max() / pose_r.abs().max()) <= POSE_RTOL
    assert int(mod.last_stats.sum()) == 0


def test_soft_fusion(cuda_device):
    ref, mod = make_pair(cuda_device, fuse_method="soft", bias_std=0.05)
    out = run_pair(ref, mod, *inputs(8, S=4, irregular=True))
    # the soft gate is a torch GEMM on the GPU vs CPU: inputs to the path differ at 1e-7
    assert out["pose_err"] <= 2 * POSE_RTOL, out
