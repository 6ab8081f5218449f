"""GPU parity of the tensor-core solver path (cfg.precision = ODEVIO_PRECISION_TF32X3, odernn_tc.cu):
per interval one cluster kernel evolves all L*B rows (3xTF32 ODEFunc GEMMs on tcgen05, solver loop in the
cluster), then the FMA kernel runs the jump + head.  Same oracle, same tolerances as the fp32 FMA path
(tests/test_odernn_gpu.py): poses <= 1e-5 max-norm relative, identical step counts wherever the reference
semantics determine them at fp32 precision."""

import pytest
import torch

from helpers import POSE_RTOL, STATE_RTOL, inputs, make_pair, rel_err, run_pair

pytestmark = pytest.mark.gpu


def _summary(out):
    return {k: v for k, v in out.items() if not isinstance(v, torch.Tensor)}


def _check(out, slack=4.0):
    s = _summary(out)
    assert out["status_max"] == 0, s
    assert out["n_mismatch_stable_entries"] <= max(1, out["n_entries"] // 200), s
    assert out["pose_err"] <= max(POSE_RTOL, slack * out["spread_pose"]), s
    assert out["h_err"] <= max(STATE_RTOL, slack * out["spread_h"]), s


def test_tc_rk4_config1(cuda_device):
    """BASELINE config 1 shape on the tensor-core path: fixed-step rk4, one (mostly padded) 128-row tile."""
    ref, mod = make_pair(cuda_device, ode_solver="rk4", ode_precision="tf32x3")
    out = run_pair(ref, mod, *inputs(16))
    _check(out)
    assert out["steps_equal"] and out["pose_err"] <= POSE_RTOL, _summary(out)


def test_tc_rk4_substeps(cuda_device):
    ref, mod = make_pair(cuda_device, ode_solver="rk4_38", ode_substeps=3, ode_precision="tf32x3", bias_std=0.05)
    out = run_pair(ref, mod, *inputs(24, S=4, irregular=True))
    _check(out)
    assert out["pose_err"] <= POSE_RTOL, _summary(out)


@pytest.mark.parametrize("solver", ["dopri5", "tsit5", "heun"])
def test_tc_adaptive_menu(cuda_device, solver):
    ref, mod = make_pair(cuda_device, ode_solver=solver, ode_precision="tf32x3", bias_std=0.05)
    out = run_pair(ref, mod, *inputs(16, irregular=True), ensemble=3)
    _check(out)


def test_tc_dopri5_irregular_rtol_ragged(cuda_device):
    """BASELINE config 2 semantics; B = 100 -> 200 rows = one full and one ragged tile, the second tile
    starts inside layer 1 (row = l * B + b)."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="tf32x3", bias_std=0.05)
    out = run_pair(ref, mod, *inputs(100, irregular=True, seed=3), ensemble=3)
    _check(out)


def test_tc_prev_state_and_gru(cuda_device):
    ref, mod = make_pair(cuda_device, ode_rnn_type="gru", ode_precision="tf32x3", bias_std=0.05)
    fv, fi, ts = inputs(16, S=5, irregular=True, offset=37.5)
    g = torch.Generator().manual_seed(5)
    prev = 0.3 * torch.randn(2, 16, 768, generator=g)
    out = run_pair(ref, mod, fv, fi, ts, prev=prev, ensemble=3)
    _check(out)


def test_tc_matches_fma_kernel_many_tiles(cuda_device):
    """More tiles than co-resident clusters (B = 1200 -> 2400 rows = 19 tiles > 16 clusters: persistent loop)
    against the fp32 FMA kernel on the same inputs: step counts equal on >= 99.5 % of entries, poses within 1e-5
    on every row with an identical step history."""
    ref, mod_tc = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="tf32x3")
    import odevio_b200
    from oracle.pose_odernn import default_opt
    mod_f = odevio_b200.PoseODERNN(default_opt(ode_solver="dopri5", ode_rtol=1e-3))
    mod_f.load_state_dict(ref.state_dict())
    mod_f = mod_f.to(cuda_device).eval()
    fv, fi, ts = (t.to(cuda_device) for t in inputs(1200, irregular=True, seed=1))
    with torch.no_grad():
        p_tc, h_tc = mod_tc(fv, fi, ts)
        p_f, h_f = mod_f(fv, fi, ts)
    torch.cuda.synchronize()
    mod_tc.check_status()
    # rows whose accept/reject history is identical in both kernels must agree to fp32 parity; a knife-edge accept
    # decision that flips (3xTF32 and FFMA differ by ~1e-7 per evaluation) changes a row by the truncation error of a
    # step (rtol = 1e-3), which is the reference semantics' own conditioning, not a kernel difference
    same = (mod_tc.last_stats == mod_f.last_stats).all(-1)            # [S, L, B]
    assert same.float().mean().item() >= 0.995, same.float().mean().item()
    rows = same.all(0).all(0).cpu()                                     # [B]
    assert rows.float().mean().item() >= 0.97, rows.float().mean().item()
    assert rel_err(p_tc.cpu()[rows], p_f.cpu()[rows]) <= POSE_RTOL
    assert rel_err(h_tc.cpu()[:, rows], h_f.cpu()[:, rows]) <= STATE_RTOL
    assert rel_err(p_tc.cpu(), p_f.cpu()) <= 1e-3


def test_tc_evolve_state(cuda_device):
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="tf32x3", bias_std=0.05)
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


def test_tc_rejects_training_and_dense(cuda_device):
    """The tensor-core path is inference-only: training falls back to the FMA kernels (precision forced to fp32
    for the checkpointed forward), the dense end-point rule is refused with a clear error."""
    import odevio_b200
    from oracle.pose_odernn import default_opt
    mod = odevio_b200.PoseODERNN(default_opt(ode_precision="tf32x3", ode_endpoint="dense")).to(cuda_device).eval()
    fv, fi, ts = (t.to(cuda_device) for t in inputs(4, S=2))
    with pytest.raises(Exception):
        with torch.no_grad():
            mod(fv, fi, ts)
