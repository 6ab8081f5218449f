"""GPU parity: fused sm_100a Neural-CDE forward (through the C ABI) vs the CPU oracle
(oracle/pose_cde.py) on the same seeded weights / inputs / timestamps.

Tolerances: poses and hidden states max-norm relative error <= 1e-5 (fp32; widened only where
noted), identical (n_steps, n_accepted, n_f_evals) -- the batch-joint controller takes one
decision per step for the whole batch, computed from an fp64-accumulated norm on both sides."""

import copy

import pytest
import torch

from helpers import rel_err
from odevio_b200 import synth
from oracle.modules import deepvio_initialization
from oracle.pose_cde import OraclePoseCDE
from oracle.pose_odernn import default_opt

pytestmark = pytest.mark.gpu

POSE_RTOL = 1e-5
HID_RTOL = 2e-5


def make_pair(dev, Hc=32, seed=0, bias_std=0.05, train=True, **over):
    import odevio_b200
    over.setdefault("cde_precision", "fp32")        # this file: the CUDA-core kernel (tests/test_cde_tc_gpu.py: tensor cores)
    opt = default_opt(v_f_len=Hc // 2, i_f_len=Hc // 2, cde_hidden_dim=Hc, **over)
    torch.manual_seed(seed)
    ref = OraclePoseCDE(opt)
    deepvio_initialization(ref)
    if bias_std > 0:
        g = torch.Generator().manual_seed(seed + 7)
        for n, p in ref.named_parameters():
            if n.endswith("bias"):
                p.data.normal_(0.0, bias_std, generator=g)
    mod = odevio_b200.PoseCDE(copy.copy(opt))
    res = mod.load_state_dict(ref.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    mod = mod.to(dev)
    ref.train(train); mod.train(train)
    return ref, mod


ULP = 2.0 ** -23


def conditioning(ref, fv, fi, ts, prev, hist, n_members=4, eps_ulps=(2.0, 8.0)):
    """How well the reference semantics pin the result at fp32 precision: re-run the ORACLE with
    2 / 8 ulp relative noise in every vector-field evaluation; returns the largest deviation of the
    members' poses from the clean fp32 oracle and whether all members reproduce its step counts."""
    with torch.no_grad():
        ref.history = hist
        p0, _ = ref(fv, fi, ts, prev=prev)
        st0 = (ref.last_stats["n_steps"], ref.last_stats["n_accepted"], ref.last_stats["n_f_evals"])
        spread, stable = 0.0, True
        for k in range(n_members):
            ref.history = hist
            ref.vf_noise = (ULP * eps_ulps[k % len(eps_ulps)], torch.Generator().manual_seed(100 + k))
            p, _ = ref(fv, fi, ts, prev=prev)
            spread = max(spread, rel_err(p, p0))
            stable &= (ref.last_stats["n_steps"], ref.last_stats["n_accepted"], ref.last_stats["n_f_evals"]) == st0
        ref.vf_noise = None
    return spread, stable


def run(ref, mod, fv, fi, ts, dev, prev=None):
    hist = None if ref.history is None else ref.history.clone()
    spread, stable = conditioning(ref, fv, fi, ts, prev, hist)
    ref.history = hist
    with torch.no_grad():
        p_ref, z_ref = ref(fv, fi, ts, prev=prev)
        p, z = mod(fv.to(dev), fi.to(dev), ts.to(dev), prev=None if prev is None else prev.to(dev))
    torch.cuda.synchronize()
    mod.check_status()
    st = mod.last_stats.cpu().tolist()
    return dict(pose_err=rel_err(p.cpu(), p_ref), z0_err=rel_err(z.cpu(), z_ref), spread=spread, stable=stable,
                stats=st, ref_stats=(ref.last_stats["n_steps"], ref.last_stats["n_accepted"], ref.last_stats["n_f_evals"]),
                pose=p.cpu(), pose_ref=p_ref)


def check(out, tol=POSE_RTOL, slack=4.0):
    """Poses within `tol`, widened ONLY to `slack` x the deviation the oracle itself shows under
    2-8 ulp noise in its vector-field evaluations; step counts identical whenever the noisy oracle
    members all reproduce them (i.e. whenever the reference semantics determine them in fp32)."""
    info = {k: v for k, v in out.items() if k not in ("pose", "pose_ref")}
    assert out["z0_err"] <= tol, info
    assert out["pose_err"] <= max(tol, slack * out["spread"]), info
    if out["stable"]:
        assert tuple(out["stats"][:3]) == tuple(out["ref_stats"]), info


def data(B, S, Hc, irregular, seed=0, offset=0.0, scale=0.2):
    """Features are scaled to 0.2: with unit-variance features the random-init CDE over a 9-knot
    horizon is ill-conditioned (the ORACLE's own fp32 and fp64 runs differ by 2e-4 in pose for
    fixed-step rk4, 3e-3 for dopri5 -- measured), so a 1e-5 comparison would test
    rounding noise, not the kernel.  At 0.2 the fp32/fp64 gap of fixed-step runs is < 3e-6."""
    fv, fi = synth.features(B, S, Hc // 2, Hc // 2, seed=seed)
    ts = synth.timestamps(B, S, irregular=irregular, seed=seed, offset=offset)
    return fv * scale, fi * scale, ts


@pytest.mark.parametrize("irregular", [False, True])
def test_reference_mode_linear_dopri5(cuda_device, irregular):
    """Reference semantics: rectilinear linear path on the integer knot grid, integrated over row
    0's times in seconds, dopri5 atol=1e-6 rtol=1e-4 (PoseCDE.py:94-101)."""
    ref, mod = make_pair(cuda_device, cde_fn_num_layers=2)
    check(run(ref, mod, *data(12, 10, 32, irregular), cuda_device))


def test_linear_crosses_knots(cuda_device):
    """Long gaps: the integration variable passes the integer knots -> steps land on jump_t, the
    vector field is re-evaluated just after each knot, segments alternate time-only / values-only."""
    ref, mod = make_pair(cuda_device, cde_fn_num_layers=2)
    fv, fi, ts = data(9, 10, 32, True, seed=3)
    ts = ts * 4.0                                  # row 0 now spans several knots
    out = run(ref, mod, fv, fi, ts, cuda_device)
    assert ref.last_stats["n_f_evals"] > 2 + 6 * ref.last_stats["n_steps"]    # jump re-evaluations happened
    check(out)


def test_cubic_dopri5(cuda_device):
    """north_star cubic control path (Hermite, backward differences), integrated over the knot grid."""
    ref, mod = make_pair(cuda_device, cde_fn_num_layers=2, cde_interp="cubic")
    out = run(ref, mod, *data(12, 10, 32, True), cuda_device)
    assert out["stats"][0] > 20
    check(out)


@pytest.mark.parametrize("interp,step", [("linear", None), ("cubic", None), ("cubic", 0.25), ("linear", 0.3)])
def test_rk4_38(cuda_device, interp, step):
    ref, mod = make_pair(cuda_device, cde_fn_num_layers=2, cde_solver="rk4", cde_interp=interp, cde_step_size=step)
    check(run(ref, mod, *data(8, 10, 32, True), cuda_device))


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "softplus"])
def test_activations(cuda_device, act):
    ref, mod = make_pair(cuda_device, cde_fn_num_layers=3, cde_activation_fn=act, cde_interp="cubic")
    check(run(ref, mod, *data(8, 6, 32, True), cuda_device))


@pytest.mark.parametrize("Hc,B,rows", [(64, 20, 0), (128, 24, 0), (128, 40, 16), (24, 7, 8)])
def test_shapes_and_tiles(cuda_device, Hc, B, rows):
    ref, mod = make_pair(cuda_device, Hc=Hc, cde_fn_num_layers=2, cde_interp="cubic", cde_rows_per_tile=rows,
                         cde_rtol=1e-3)
    check(run(ref, mod, *data(B, 5, Hc, True), cuda_device))


def test_many_tiles_per_cta(cuda_device):
    """More tiles than SMs: every CTA integrates several tiles in lock-step under one joint controller."""
    ref, mod = make_pair(cuda_device, Hc=16, cde_fn_num_layers=1, cde_interp="cubic", cde_rows_per_tile=8,
                         cde_rtol=1e-3)
    check(run(ref, mod, *data(8 * 148 + 13, 3, 16, True), cuda_device))


def test_eval_mode_history_and_prev(cuda_device):
    """eval mode: absolute timestamps, history grows across windows, z0 = prev (PoseCDE.py:81,88-96)."""
    ref, mod = make_pair(cuda_device, cde_fn_num_layers=2, train=False)
    fv, fi, ts = data(6, 8, 32, True, seed=5)
    o1 = run(ref, mod, fv[:, :4], fi[:, :4], ts[:, :5], cuda_device)
    check(o1)
    g = torch.Generator().manual_seed(1)
    prev = 0.3 * torch.randn(6, 32, generator=g)
    o2 = run(ref, mod, fv[:, 4:], fi[:, 4:], ts[:, 4:], cuda_device, prev=prev)
    assert ref.history.shape[1] == 8 and mod.history[0].shape[1] == 8
    check(o2)


def test_joint_controller_depends_on_batch(cuda_device):
    """torchdiffeq's step control is batch-joint: the same row integrated in a different batch takes
    different steps (documented consequence: CDE batches are not shard-invariant)."""
    ref, mod = make_pair(cuda_device, cde_fn_num_layers=2, cde_interp="cubic")
    fv, fi, ts = data(16, 6, 32, True)
    a = run(ref, mod, fv, fi, ts, cuda_device)
    b = run(ref, mod, fv[:4], fi[:4], ts[:4], cuda_device)
    check(a); check(b)
    assert a["stats"][0] != b["stats"][0] or not torch.equal(a["pose"][:4], b["pose"])


@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_configs2_full_size(cuda_device, interp):
    """BASELINE configs[2] at FULL size: B = 1024, Hc = F = 128, n = 3, dopri5 atol=1e-6 rtol=1e-4, irregular
    timestamps; reference rectilinear-linear path and north_star cubic path.  The batch-joint controller
    cannot be checked on a row subset (one step size for the whole batch), so the CPU oracle integrates all
    1024 sequences (~2 TFLOP for the cubic path: seconds).  Same criterion as the small cases: poses <= 1e-5
    (widened only to 4x the oracle's own 2-8 ulp noise spread), identical (n_steps, n_accepted, n_f_evals)
    whenever the noisy oracle members reproduce them."""
    ref, mod = make_pair(cuda_device, Hc=128, cde_fn_num_layers=3, cde_interp=interp)
    fv, fi, ts = data(1024, 10, 128, True)
    hist = None
    spread, stable = conditioning(ref, fv, fi, ts, None, hist, n_members=2)
    with torch.no_grad():
        p_ref, z_ref = ref(fv, fi, ts)
        p, z = mod(fv.to(cuda_device), fi.to(cuda_device), ts.to(cuda_device))
    torch.cuda.synchronize()
    mod.check_status()
    out = dict(pose_err=rel_err(p.cpu(), p_ref), z0_err=rel_err(z.cpu(), z_ref), spread=spread, stable=stable,
               stats=mod.last_stats.cpu().tolist(),
               ref_stats=(ref.last_stats["n_steps"], ref.last_stats["n_accepted"], ref.last_stats["n_f_evals"]))
    print(f"configs[2] {interp}: pose_err {out['pose_err']:.3e} (oracle noise spread {spread:.3e}, stable={stable}), "
          f"stats kernel {out['stats'][:3]} oracle {out['ref_stats']}")
    check(out)


@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_reference_run_shape_hc400(cuda_device, interp):
    """The reference's one PoseCDE run (scripts/run_training.sh:61-70) uses cde_hidden_dim = 400: a 400 x 401 x 400 final Linear
    (257 MB, beyond L2; Gc = 2 channels per group, 201 groups).  Small batch: the CUDA-core kernel against the oracle, same
    criterion as everywhere; the tensor-core kernel does not take this width ("auto" falls back, asserted)."""
    import time
    ref, mod = make_pair(cuda_device, Hc=400, cde_fn_num_layers=2, cde_interp=interp, cde_rtol=1e-3, cde_precision="auto")
    fv, fi, ts = data(6, 4, 400, True)
    out = run(ref, mod, fv, fi, ts, cuda_device)
    assert mod.last_precision == "fp32"
    with torch.no_grad():
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mod(fv.to(cuda_device), fi.to(cuda_device), ts.to(cuda_device))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"Hc = 400 {interp}: pose_err {out['pose_err']:.2e}, stats {out['stats'][:3]}, {dt * 1e3:.1f} ms per forward (B = 6)")
    check(out)


def test_unit_variance_features_measured_spread(cuda_device):
    """BASELINE's N(0,1) features (the other CDE tests and the bench scale them to 0.2, see data()).
    With unit-variance features the random-init cubic CDE is ill-conditioned: this test MEASURES and prints
    the oracle's own fp32-vs-fp64 pose gap and its 2-8 ulp noise spread, and holds the kernel to 4x the
    larger of the two (never tighter than 1e-5) -- i.e. the kernel must be as close to the fp32 oracle as
    the fp32 oracle is pinned by the reference semantics at this conditioning."""
    ref, mod = make_pair(cuda_device, Hc=32, cde_fn_num_layers=2, cde_interp="cubic")
    fv, fi, ts = data(12, 10, 32, True, scale=1.0)
    spread, stable = conditioning(ref, fv, fi, ts, None, None)
    ref64 = copy.deepcopy(ref).double()
    with torch.no_grad():
        p_ref, _ = ref(fv, fi, ts)
        p64, _ = ref64(fv.double(), fi.double(), ts.double())
        p, _ = mod(fv.to(cuda_device), fi.to(cuda_device), ts.to(cuda_device))
    mod.check_status()
    gap64 = rel_err(p_ref.double(), p64)
    err = rel_err(p.cpu(), p_ref)
    print(f"unit-variance features: kernel vs fp32 oracle {err:.3e}; oracle fp32 vs fp64 {gap64:.3e}; "
          f"oracle 2-8 ulp noise spread {spread:.3e}; steps stable under noise: {stable}")
    assert err <= max(POSE_RTOL, 4 * max(spread, gap64))


def test_bounded_history_windows_match_unbounded_oracle(cuda_device):
    """f2: chained eval-mode windows with `cde_history_limit` -- the module keeps max(limit, S + 1)
    observations instead of the reference's ever-growing history (PoseCDE.py:88-92) -- against the ORACLE running the
    reference's unbounded history.  Cubic mode; the two paths are the same function, the solutions agree within the solver
    tolerance (the knot index of the unbounded run loses ulp(t) of the in-segment parameter: tests/test_oracle_cde.py)."""
    ref, mod = make_pair(cuda_device, Hc=32, cde_fn_num_layers=2, cde_interp="cubic", train=False, cde_history_limit=1)
    ref.history_limit = None                      # the oracle keeps everything, as the reference does
    ref.eval(); mod.eval()
    g = torch.Generator().manual_seed(5)
    B, S = 9, 5
    t0 = torch.zeros(B, 1)
    hc_ref = hc = None
    for w in range(4):
        fv, fi = 0.2 * torch.randn(B, S, 16, generator=g), 0.2 * torch.randn(B, S, 16, generator=g)
        ts = torch.cat([t0, t0 + torch.cumsum(0.1 + 0.1 * torch.rand(B, S, generator=g), 1)], 1)
        t0 = ts[:, -1:]
        with torch.no_grad():
            p_ref, hc_ref = ref(fv, fi, ts, prev=hc_ref)
            p, hc = mod(fv.to(cuda_device), fi.to(cuda_device), ts.to(cuda_device), prev=hc)
        mod.check_status()
        assert ref.history.shape[1] == (w + 1) * S and mod.history[0].shape[1] == min((w + 1) * S, S + 1)
        assert rel_err(p.cpu(), p_ref) <= 2e-4, (w, rel_err(p.cpu(), p_ref))
