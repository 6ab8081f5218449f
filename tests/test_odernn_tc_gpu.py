"""GPU parity of the tensor-core solver paths.  cfg.precision = ODEVIO_PRECISION_TF32X3 (odernn_tc.cu): per interval one
cluster kernel evolves all L*B rows, then the FMA kernel runs the jump + head.  ODEVIO_PRECISION_FP16X3 (odernn_h3.cu):
ONE launch per forward -- solver loops of all intervals, rnn jump and pose head inside the cluster kernel (tanh rnn, "cat"
fusion, L <= 2; GRU / "soft" fusion: per-interval launches + FMA jump).  Same oracle, same tolerances as the fp32 FMA path
(tests/test_odernn_gpu.py): poses <= 1e-5 max-norm relative, identical step counts wherever the reference
semantics determine them at fp32 precision."""

import pytest
import torch

from helpers import POSE_RTOL, STATE_RTOL, inputs, make_pair, rel_err, run_pair

pytestmark = pytest.mark.gpu

# both tensor-core solvers behind the same ABI switch: "tf32x3" = odernn_tc.cu (clusters of 8 / 4, 128-row tiles, 3xTF32),
# "fp16x3" = odernn_h3.cu (clusters of 4, 64-row tiles, weights on the M side, 3xFP16; the bench default)
PRECISIONS = ["tf32x3", "fp16x3"]


def _summary(out):
    return {k: v for k, v in out.items() if not isinstance(v, torch.Tensor)}


def _check(out, slack=4.0):
    s = _summary(out)
    assert out["status_max"] == 0, s
    assert out["n_mismatch_stable_entries"] <= max(1, out["n_entries"] // 200), s
    assert out["pose_err"] <= max(POSE_RTOL, slack * out["spread_pose"]), s
    assert out["h_err"] <= max(STATE_RTOL, slack * out["spread_h"]), s


@pytest.mark.parametrize("prec", PRECISIONS)
def test_tc_rk4_config1(cuda_device, prec):
    """BASELINE config 1 shape on the tensor-core path: fixed-step rk4, one (mostly padded) 128-row tile."""
    ref, mod = make_pair(cuda_device, ode_solver="rk4", ode_precision=prec)
    out = run_pair(ref, mod, *inputs(16))
    _check(out)
    assert out["steps_equal"] and out["pose_err"] <= POSE_RTOL, _summary(out)


@pytest.mark.parametrize("prec", PRECISIONS)
def test_tc_rk4_substeps(cuda_device, prec):
    ref, mod = make_pair(cuda_device, ode_solver="rk4_38", ode_substeps=3, ode_precision=prec, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(24, S=4, irregular=True))
    _check(out)
    assert out["pose_err"] <= POSE_RTOL, _summary(out)


@pytest.mark.parametrize("solver", ["dopri5", "tsit5", "heun"])
@pytest.mark.parametrize("prec", PRECISIONS)
def test_tc_adaptive_menu(cuda_device, solver, prec):
    ref, mod = make_pair(cuda_device, ode_solver=solver, ode_precision=prec, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(16, irregular=True), ensemble=3)
    _check(out)


@pytest.mark.parametrize("prec", PRECISIONS)
def test_tc_dopri5_irregular_rtol_ragged(cuda_device, prec):
    """BASELINE config 2 semantics; B = 100 -> 200 rows = one full and one ragged tile, the second tile
    starts inside layer 1 (row = l * B + b)."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision=prec, bias_std=0.05)
    out = run_pair(ref, mod, *inputs(100, irregular=True, seed=3), ensemble=3)
    _check(out)


@pytest.mark.parametrize("prec", PRECISIONS)
def test_tc_prev_state_and_gru(cuda_device, prec):
    ref, mod = make_pair(cuda_device, ode_rnn_type="gru", ode_precision=prec, bias_std=0.05)
    fv, fi, ts = inputs(16, S=5, irregular=True, offset=37.5)
    g = torch.Generator().manual_seed(5)
    prev = 0.3 * torch.randn(2, 16, 768, generator=g)
    out = run_pair(ref, mod, fv, fi, ts, prev=prev, ensemble=3)
    _check(out)


@pytest.mark.parametrize("B", [1024, 1200, 2600])
@pytest.mark.parametrize("prec", PRECISIONS)
def test_tc_full_size_rows_match_oracle_subset(cuda_device, B, prec):
    """configs[1] at full size on the tensor-core path.  B = 1024 -> 2048 rows = 16 tiles: on a GPU that holds 15
    clusters 64 sequences per interval run concurrently in the FMA kernel (side launch); B = 1200 -> 19 tiles: the
    clusters-of-4 pre-split instantiation, one round; B = 2600 -> 41 tiles > 37 co-resident clusters of 4: its
    persistent loop takes a second round.  The oracle runs a random subset of rows (rows are independent);
    same criterion as tests/test_full_size_gpu.py.  Also: the rows agree with the fp32 FMA kernel to 1e-5 wherever
    both kernels took the same accept/reject history."""
    import odevio_b200
    from helpers import noise_ensemble
    from oracle.pose_odernn import default_opt
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision=prec, bias_std=0.05)
    fv, fi, ts = inputs(B, 10, irregular=True, seed=0)
    dev = cuda_device
    with torch.no_grad():
        p, h = mod(fv.to(dev), fi.to(dev), ts.to(dev))
    assert int(mod.last_status.max().item()) == 0
    rows = torch.cat([torch.randperm(B, generator=torch.Generator().manual_seed(3))[:20],
                      torch.tensor([0, B // 2, B - 2, B - 1])])        # incl. the last rows (side launch / ragged tile)
    with torch.no_grad():
        p_ref, h_ref = ref(fv[rows], fi[rows], ts[rows])
    st = mod.last_stats.cpu().long()[:, :, rows]
    neq = (st[..., 0] != ref.last_stats["n_steps"]) | (st[..., 1] != ref.last_stats["n_accepted"])
    stable, spread_p, spread_h = noise_ensemble(ref, fv[rows], fi[rows], ts[rows], n_members=6)
    assert int((neq & stable).sum()) <= max(1, neq.numel() // 200), (int(neq.sum()), int((~stable).sum()))
    assert rel_err(p.cpu()[rows], p_ref) <= max(POSE_RTOL, 4 * spread_p), (rel_err(p.cpu()[rows], p_ref), spread_p)
    assert rel_err(h.cpu()[:, rows], h_ref) <= max(STATE_RTOL, 4 * spread_h), (rel_err(h.cpu()[:, rows], h_ref), spread_h)
    # against the FMA kernel, all rows
    mod_f = odevio_b200.PoseODERNN(default_opt(ode_solver="dopri5", ode_rtol=1e-3))
    mod_f.load_state_dict(ref.state_dict())
    mod_f = mod_f.to(dev).eval()
    with torch.no_grad():
        p_f, h_f = mod_f(fv.to(dev), fi.to(dev), ts.to(dev))
    same_rows = (mod.last_stats == mod_f.last_stats).all(-1).all(0).all(0).cpu()
    # the step counts themselves are ill-conditioned (helpers.noise_ensemble flags ~2/3 of the entries), so only a
    # minority of rows shares the complete 20-entry history; those must agree to fp32 parity
    assert int(same_rows.sum()) >= 16, int(same_rows.sum())
    # both kernels are within 1e-5 of the oracle, so within 2e-5 of each other
    cross = rel_err(p.cpu()[same_rows], p_f.cpu()[same_rows])
    assert cross <= 2 * POSE_RTOL, cross
    assert rel_err(p.cpu(), p_f.cpu()) <= 1e-3


@pytest.mark.parametrize("prec", PRECISIONS)
def test_tc_evolve_state(cuda_device, prec):
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision=prec, bias_std=0.05)
    g = torch.Generator().manual_seed(11)
    B = 37
    state = 0.5 * torch.randn(B, mod.f_len, generator=g)
    t0 = torch.rand(B, generator=g)
    ts = torch.stack([t0, t0 + 0.1 + 0.3 * torch.rand(B, generator=g)], 1)
    with torch.no_grad():
        want = ref.evolve_state(state, ts)
        got = mod.evolve_state(state.to(cuda_device), ts.to(cuda_device))
    mod.check_status()
    assert rel_err(got.cpu(), want["y_end"]) <= STATE_RTOL
    ns = mod.last_stats[0, 0, :, 0].cpu().long()
    assert (ns == want["n_steps"]).float().mean().item() >= 0.97


def test_h3_literal_torchode_end_point(cuda_device):
    """The literal torchode end point -- quartic dense output of the step that reaches t_end, no exact landing (SURVEY.md
    A.1's reading of `ode_solution.ys[:, -1, :]`, PoseODERNN.py:75) -- on the one-launch tcgen05 kernel, held to the bound
    the FMA kernel's test uses (tests/test_odernn_gpu.py::test_literal_torchode_arithmetic: the mode is ill-conditioned in
    fp32, the ORACLE itself moves by > 1e-5 under 1e-7 weight noise); plus the well-conditioned variant (dense end point
    WITH exact landing: x = 1, the quartic collapses onto y1 up to rounding) at the usual 1e-5 criterion."""
    ref, mod = make_pair(cuda_device, ode_endpoint="dense", ode_exact_landing=False, bias_std=0.05, ode_precision="fp16x3")
    out = run_pair(ref, mod, *inputs(16, irregular=True))
    assert out["status_max"] == 0
    print(f"literal end point on fp16x3: pose {out['pose_err']:.2e}, h {out['h_err']:.2e}")
    assert out["pose_err"] <= 5e-4, _summary(out)
    assert out["h_err"] <= 5e-4, _summary(out)
    ref, mod = make_pair(cuda_device, ode_endpoint="dense", ode_exact_landing=True, bias_std=0.05, ode_precision="fp16x3")
    out = run_pair(ref, mod, *inputs(16, irregular=True), ensemble=3)
    _check(out)
    # heun has no b_mid: linear interpolant between y0 and y1
    ref, mod = make_pair(cuda_device, ode_endpoint="dense", ode_exact_landing=False, bias_std=0.05, ode_precision="fp16x3",
                         ode_solver="heun", ode_rtol=1e-3)
    out = run_pair(ref, mod, *inputs(8, S=4, irregular=True))
    assert out["status_max"] == 0 and out["pose_err"] <= 5e-4, _summary(out)


def test_tc_rejects_dense_on_tf32x3(cuda_device, prec="tf32x3"):
    """The round-1 3xTF32 kernel has no dense end-point rule: refused with a clear error (the FP16X3 kernel has it)."""
    import odevio_b200
    from oracle.pose_odernn import default_opt
    mod = odevio_b200.PoseODERNN(default_opt(ode_precision=prec, ode_endpoint="dense")).to(cuda_device).eval()
    fv, fi, ts = (t.to(cuda_device) for t in inputs(4, S=2))
    with pytest.raises(Exception):
        with torch.no_grad():
            mod(fv, fi, ts)


@pytest.mark.parametrize("prec", PRECISIONS)
def test_tc_evolve_state_side_launch_l1(cuda_device, prec):
    """evolve_state is an L = 1 problem: 2048 rows = 16 tiles -> with 15 co-resident clusters the 128 shortest-interval
    rows run in the FFMA side launch as 8-row `<8, 1>` tiles.  Fixed-step rk4 (no controller feedback), so every row must
    agree with the FFMA-only kernel to fp32 parity, whichever kernel integrated it."""
    import odevio_b200
    from oracle.pose_odernn import default_opt
    ref, mod = make_pair(cuda_device, ode_solver="rk4", ode_substeps=2, ode_precision=prec, bias_std=0.05)
    mod_f = odevio_b200.PoseODERNN(default_opt(ode_solver="rk4", ode_substeps=2))
    mod_f.load_state_dict(ref.state_dict())
    mod_f = mod_f.to(cuda_device).eval()
    g = torch.Generator().manual_seed(4)
    B = 2048
    state = (0.5 * torch.randn(B, mod.f_len, generator=g)).to(cuda_device)
    t0 = torch.rand(B, generator=g)
    ts = torch.stack([t0, t0 + 0.05 + 0.3 * torch.rand(B, generator=g)], 1).to(cuda_device)
    with torch.no_grad():
        got = mod.evolve_state(state, ts)
        want = mod_f.evolve_state(state, ts)
    mod.check_status()
    assert rel_err(got.cpu(), want.cpu()) <= STATE_RTOL
    assert (mod.last_stats[0, 0, :, 0] == 2).all() and torch.equal(mod.last_stats, mod_f.last_stats)
    rows = torch.randperm(B, generator=g)[:16]
    with torch.no_grad():
        o = ref.evolve_state(state.cpu()[rows], ts.cpu()[rows])
    assert rel_err(got.cpu()[rows], o["y_end"]) <= STATE_RTOL


def test_h3_single_launch_l1_prev_and_zero_length_intervals(cuda_device):
    """The one-launch forward of odernn_h3.cu at L = 1 (64 sequences per tile, jump MMAs over 64 rows), with a carried
    state `prev`, a ragged last tile (B = 70) and observation intervals of length zero for some rows (no solve, the jump
    still runs: PoseODERNN.py:108-117)."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="fp16x3", rnn_num_layers=1, bias_std=0.05)
    fv, fi, ts = inputs(70, S=6, irregular=True, seed=5, offset=12.0)
    ts[::3, 3] = ts[::3, 2]                      # zero-length third interval on every third sequence
    ts[5, 1] = ts[5, 0]
    g = torch.Generator().manual_seed(9)
    prev = 0.3 * torch.randn(1, 70, 768, generator=g)
    out = run_pair(ref, mod, fv, fi, ts, prev=prev, ensemble=3)
    _check(out)
    st = mod.last_stats.cpu()
    assert (st[2, 0, ::3, 0] == 0).all() and st[0, 0, 5, 0] == 0


def test_h3_single_launch_prev_l2(cuda_device):
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="fp16x3", bias_std=0.05)
    fv, fi, ts = inputs(45, S=4, irregular=True, seed=2, offset=3.25)
    g = torch.Generator().manual_seed(6)
    prev = 0.3 * torch.randn(2, 45, 768, generator=g)
    out = run_pair(ref, mod, fv, fi, ts, prev=prev, ensemble=3)
    _check(out)


def test_h3_soft_fusion_takes_the_per_interval_path(cuda_device):
    """FusionModule "soft" is evaluated in the FMA kernel's jump prologue: the fp16x3 mode then launches per interval."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="fp16x3", fuse_method="soft", bias_std=0.05)
    out = run_pair(ref, mod, *inputs(20, S=3, irregular=True, seed=4), ensemble=3)
    _check(out)


def test_h3_weights_prepacked_reuse(cuda_device):
    """Second forward with unchanged weights skips the packing launches (cfg.weights_prepacked) and is bit-identical;
    an in-place weight update invalidates the cache."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="fp16x3", bias_std=0.05)
    fv, fi, ts = (t.to(cuda_device) for t in inputs(40, S=3, irregular=True, seed=1))
    with torch.no_grad():
        p1, _ = mod(fv, fi, ts)
        assert not mod.last_prepacked
        p2, _ = mod(fv, fi, ts)
        assert mod.last_prepacked and torch.equal(p1, p2)
        mod.regressor[2].bias.add_(1.0)
        p3, _ = mod(fv, fi, ts)
        assert not mod.last_prepacked
    assert torch.allclose(p3, p1 + 1.0, atol=1e-5)
