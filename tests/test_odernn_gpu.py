"""GPU parity: fused sm_100a ODE-RNN forward (through the C ABI) vs the CPU oracle on the same
seeded weights / inputs / timestamps.  Tolerance: max-norm relative pose error <= 1e-5 (fp32),
identical per-row step counts (n_steps, n_accepted) for the adaptive solvers."""

import pytest
import torch

from helpers import POSE_RTOL, STATE_RTOL, inputs, make_pair, run_pair

pytestmark = pytest.mark.gpu


def _summary(out):
    return {k: v for k, v in out.items() if not isinstance(v, torch.Tensor)}


def _check(out, pose_tol=POSE_RTOL, state_tol=STATE_RTOL, slack=4.0):
    """Tolerances: max-norm relative pose error <= 1e-5 (north_star), hidden state <= 5e-5 --
    widened ONLY to `slack` x the deviation the oracle itself shows under 1-ulp noise in its
    vector-field evaluations (helpers.noise_ensemble; 0 for fixed-step solvers, where the strict
    bound applies).  Step counts must be identical on every entry the reference semantics
    determine at fp32 precision."""
    s = _summary(out)
    assert out["status_max"] == 0, s
    # the Monte-Carlo ensemble is finite: allow one knife-edge entry (or 0.5 %) it did not flag
    assert out["n_mismatch_stable_entries"] <= max(1, out["n_entries"] // 200), s
    assert out["pose_err"] <= max(pose_tol, slack * out["spread_pose"]), s
    assert out["h_err"] <= max(state_tol, slack * out["spread_h"]), s


def test_config1_rk4(cuda_device):
    """BASELINE config 1: fixed-step rk4, B=16, seq_len 11."""
    ref, mod = make_pair(cuda_device, ode_solver="rk4")
    out = run_pair(ref, mod, *inputs(16))
    _check(out)
    assert out["steps_equal"] and out["pose_err"] <= POSE_RTOL     # strict: no controller feedback


def test_reference_init_dopri5_strict(cuda_device):
    """Reference init (zero biases, DeepVIO.py:77-87), regular 10 Hz frames, reference tolerances:
    the strict value bounds apply -- 1e-5 poses -- and step counts agree wherever determined."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", bias_std=0.0)
    out = run_pair(ref, mod, *inputs(16), ensemble=3)
    assert out["status_max"] == 0 and out["n_mismatch_stable_entries"] <= 1, _summary(out)
    assert out["pose_err"] <= POSE_RTOL and out["h_err"] <= STATE_RTOL, _summary(out)


@pytest.mark.parametrize("solver", ["rk4_38", "dopri5", "tsit5", "heun"])
def test_solver_menu(cuda_device, solver):
    ref, mod = make_pair(cuda_device, ode_solver=solver, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(16, irregular=True), ensemble=3)
    _check(out)


def test_euler_fixed_dt(cuda_device):
    # torchode Euler keeps dt = dt0 -> 1e-4 steps would be 1000 steps/interval; use a coarser dt0
    ref, mod = make_pair(cuda_device, ode_solver="euler", ode_dt0=0.02, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(8, S=4))
    _check(out)


def test_dopri5_irregular_rtol(cuda_device):
    """BASELINE config 2 semantics at a size the oracle finishes quickly."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(64, irregular=True, seed=3), ensemble=3)
    _check(out)


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "softplus"])
def test_activations(cuda_device, act):
    ref, mod = make_pair(cuda_device, ode_activation_fn=act, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(8, S=5, irregular=True), ensemble=3)
    _check(out)


def test_gru_jump(cuda_device):
    ref, mod = make_pair(cuda_device, ode_rnn_type="gru", bias_std=0.05)
    out = run_pair(ref, mod, *inputs(16, irregular=True), ensemble=3)
    _check(out)


@pytest.mark.parametrize("L,H,n", [(1, 128, 2), (3, 1024, 2), (3, 512, 3), (4, 256, 1), (2, 512, 4), (2, 1024, 3)])
def test_shapes(cuda_device, L, H, n):
    ref, mod = make_pair(cuda_device, rnn_num_layers=L, ode_hidden_dim=H, ode_fn_num_layers=n, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(11, S=4, irregular=True), ensemble=3)      # B not a multiple of the tile
    _check(out)


@pytest.mark.parametrize("rt", [4, 8, 16])
def test_tile_heights(cuda_device, rt):
    ref, mod = make_pair(cuda_device, ode_rows_per_tile=rt, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(24, S=4, irregular=True), ensemble=3)
    _check(out)


def test_literal_torchode_arithmetic(cuda_device):
    """Literal fp32 torchode arithmetic (quartic dense output at x=1, t + (t_end - t) landing) is
    ill-conditioned: tests/test_oracle.py shows the ORACLE moves by > 1e-5 in pose and changes step
    counts under 1e-7 weight noise.  The kernel still implements it; it is checked at the looser
    bound that conditioning allows."""
    ref, mod = make_pair(cuda_device, ode_endpoint="dense", ode_exact_landing=False, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(16, irregular=True))
    assert out["status_max"] == 0
    assert out["pose_err"] <= 5e-4, out
    assert out["h_err"] <= 5e-4, out


def test_prev_carry_absolute_time(cuda_device):
    """hc carried across windows with ABSOLUTE timestamps (KITTI_eval.py:141, PoseODERNN.py:97-100)."""
    ref, mod = make_pair(cuda_device, bias_std=0.05)
    fv, fi, ts = inputs(8, S=6, irregular=True, offset=123.0)
    g = torch.Generator().manual_seed(5)
    prev = 0.5 * torch.randn(2, 8, 768, generator=g)
    out = run_pair(ref, mod, fv, fi, ts, prev=prev, ensemble=3)
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
    out = run_pair(ref, mod, *inputs(8, S=4, irregular=True), ensemble=3)
    _check(out)                          # the gate cat * Linear(cat) is evaluated in the kernel prologue
    ref2, mod2 = make_pair(cuda_device, fuse_method="soft", ode_rnn_type="gru", ode_solver="rk4", bias_std=0.05)
    _check(run_pair(ref2, mod2, *inputs(11, S=3)))


def test_controller_trace_matches_oracle(cuda_device):
    """Direct check of the in-kernel error norm + step controller through the (dt, ratio) trace:
    wherever kernel and oracle took the same step size and the error estimate is above the
    rounding-noise floor (ratio > 1e-2; the noise floor is ~1e-4 at these step sizes), the error ratios agree to 3 %; the first three step
    sizes of every solve are the deterministic 1e-4, 1e-3, 1e-2 ramp (factor clamped at 10)."""
    ref, mod = make_pair(cuda_device, bias_std=0.05, ode_trace_steps=6, ode_rtol=1e-3)
    fv, fi, ts = inputs(32, irregular=True, seed=11)
    run_pair(ref, mod, fv, fi, ts)
    tg, tr = mod.last_trace.cpu(), ref.last_stats["trace"]
    took = (tg[..., 0] > 0) & (tr[..., 0] > 0)
    same_dt = took & ((tg[..., 0] - tr[..., 0]).abs() <= 1e-6 * tr[..., 0])
    sig = same_dt & (tr[..., 1] > 1e-2)
    assert sig.sum() > 5
    rel = ((tg[..., 1] - tr[..., 1]).abs() / tr[..., 1])[sig]
    assert rel.max() <= 3e-2, (rel.max(), rel.median(), int(sig.sum()))
    # interval >= 1 (state away from 0): ramp-up is noise-free on both sides
    ramp = tg[1:, :, :, :2, 0]
    assert torch.allclose(ramp, tr[1:, :, :, :2, 0], rtol=1e-6, atol=0)


@pytest.mark.parametrize("solver", ["dopri5", "rk4"])
def test_evolve_state(cuda_device, solver):
    """PoseODERNN.evolve_state (reference PoseODERNN.py:70-75): one IVP per row, no jump / head
    (kernel flag evolve_only).  Also checks that two chained calls equal one call over three
    time columns (consecutive intervals restart the controller at dt0, like the reference loop)."""
    ref, mod = make_pair(cuda_device, ode_solver=solver, ode_rtol=1e-3, bias_std=0.05)
    g = torch.Generator().manual_seed(11)
    B = 37                                              # ragged: not a multiple of any tile
    state = 0.5 * torch.randn(B, mod.f_len, generator=g)
    t0 = torch.rand(B, generator=g)
    ts = torch.stack([t0, t0 + 0.1 + 0.3 * torch.rand(B, generator=g)], 1)
    with torch.no_grad():
        want = ref.evolve_state(state, ts)
        got = mod.evolve_state(state.to(cuda_device), ts.to(cuda_device))
    mod.check_status()
    err = ((got.cpu() - want["y_end"]).abs().max() / want["y_end"].abs().max()).item()
    assert err <= STATE_RTOL, err
    ns = mod.last_stats[0, 0, :, 0].cpu().long()
    assert (ns == want["n_steps"]).float().mean().item() >= 0.97, (ns, want["n_steps"])
    ts3 = torch.cat([ts, ts[:, 1:] + 0.2], 1)
    with torch.no_grad():
        both = mod.evolve_state(state.to(cuda_device), ts3.to(cuda_device))
        second = mod.evolve_state(got, ts3[:, 1:].to(cuda_device))
    assert torch.equal(both, second)
