"""Full-size checks (BASELINE.json configs): the oracle cannot run B = 1024 in seconds, so parity at
full size is established through (a) the oracle on a random SUBSET of rows -- rows are independent,
so the fused kernel's rows of a 1024-sequence launch must equal the oracle run on those rows alone --
and (b) size-independent properties: shard invariance, window chaining, gradient linearity."""

import pytest
import torch

from helpers import POSE_RTOL, inputs, make_pair, noise_ensemble, rel_err

pytestmark = pytest.mark.gpu


def test_config2_full_batch_rows_match_oracle_subset(cuda_device):
    """configs[1]: dopri5 rtol=1e-3, irregular timestamps, B = 1024."""
    ref, mod = make_pair(cuda_device, ode_solver="dopri5", ode_rtol=1e-3, bias_std=0.05)
    fv, fi, ts = inputs(1024, 10, irregular=True, seed=0)
    dev = cuda_device
    with torch.no_grad():
        p, h = mod(fv.to(dev), fi.to(dev), ts.to(dev))
    assert int(mod.last_status.max().item()) == 0
    rows = torch.randperm(1024, generator=torch.Generator().manual_seed(3))[:24]
    with torch.no_grad():
        p_ref, h_ref = ref(fv[rows], fi[rows], ts[rows])
    st = mod.last_stats.cpu().long()[:, :, rows]
    neq = (st[..., 0] != ref.last_stats["n_steps"]) | (st[..., 1] != ref.last_stats["n_accepted"])
    # same criterion as tests/test_odernn_gpu.py: step counts identical wherever the oracle itself is
    # stable under 2-32 ulp noise; poses within 1e-5, widened only to 4x the oracle's own noise spread
    stable, spread_p, _ = noise_ensemble(ref, fv[rows], fi[rows], ts[rows], n_members=6)
    # raw numbers, un-widened (pytest -s / the GPU log): what the 1e-5 criterion sees before the noise allowance
    print(f"configs[1] B=1024, {len(rows)} rows vs oracle: pose err {rel_err(p.cpu()[rows], p_ref):.3e} (oracle noise spread "
          f"{spread_p:.3e}), step-count mismatches {int(neq.sum())}/{neq.numel()} of which {int((neq & stable).sum())} on entries "
          f"the oracle's own 2-32 ulp ensemble determines ({int((~stable).sum())} undetermined)")
    assert int((neq & stable).sum()) <= max(1, neq.numel() // 200), (int(neq.sum()), int((~stable).sum()))
    assert rel_err(p.cpu()[rows], p_ref) <= max(POSE_RTOL, 4 * spread_p), (rel_err(p.cpu()[rows], p_ref), spread_p)
    # shard invariance at full size: bit-identical rows
    with torch.no_grad():
        p1, h1 = mod(fv[:512].to(dev), fi[:512].to(dev), ts[:512].to(dev))
        p2, h2 = mod(fv[512:].to(dev), fi[512:].to(dev), ts[512:].to(dev))
    assert torch.equal(torch.cat([p1, p2], 0), p) and torch.equal(torch.cat([h1, h2], 1), h)


def test_streaming_windows_equal_one_call(cuda_device):
    """Chained windows with the carried state == one long call on absolute times (KITTI_eval.py:124-160)."""
    from odevio_b200.streaming import StreamingPoseODERNN
    ref, mod = make_pair(cuda_device, bias_std=0.05)
    fv, fi, ts = inputs(64, 30, irregular=True, seed=4, offset=12.0)
    dev = cuda_device
    h0 = torch.zeros(2, 64, 768, device=dev)
    with torch.no_grad():
        p_all, h_all = mod(fv.to(dev), fi.to(dev), ts.to(dev), prev=h0)
    stream = StreamingPoseODERNN(mod)
    stream.state = h0.clone()
    p_chain = stream.run(fv.to(dev), fi.to(dev), ts.to(dev), window=10)
    assert torch.equal(p_chain, p_all) and torch.equal(stream.state, h_all)


def test_training_gradient_is_linear_in_shards(cuda_device):
    """configs[3] property: the mean-loss gradient of a batch is the row-weighted sum of its shards'
    gradients (what the NCCL all-reduce relies on), here 192 = 128 + 64 rows through the fused backward."""
    from odevio_b200.distributed import pose_loss
    ref, mod = make_pair(cuda_device, bias_std=0.05, ode_rtol=1e-3)
    mod.train()
    dev = cuda_device
    fv, fi, ts = inputs(192, 10, irregular=True, seed=6)
    gts = 0.05 * torch.randn(192, 10, 6, generator=torch.Generator().manual_seed(2))

    def grads(a, b):
        mod.zero_grad(set_to_none=True)
        p, _ = mod(fv[a:b].to(dev), fi[a:b].to(dev), ts[a:b].to(dev))
        pose_loss(p, gts[a:b].to(dev)).backward()
        return {n: q.grad.clone() for n, q in mod.named_parameters() if q.grad is not None}

    g_all, g1, g2 = grads(0, 192), grads(0, 128), grads(128, 192)
    for n in g_all:
        comb = g1[n] * (128 / 192) + g2[n] * (64 / 192)
        assert rel_err(comb, g_all[n]) <= 2e-5, n


def test_config4_full_batch_gradient_rows_match_oracle_subset(cuda_device):
    """configs[3] at FULL size: training step at B = 4096 (dopri5 rtol=1e-3, irregular timestamps) through
    the fused forward + backward.  Rows are independent, and d loss / d features of a row depends on that
    row only, so the fused backward's input gradients of a random 16-row subset must equal autograd through
    the CPU oracle run on those 16 rows alone with the loss scaled by 16 / 4096 (mean reduction over the
    batch, scripts/train_model.py:72-77).  The parameter gradient (a sum over rows) is covered by
    test_training_gradient_is_linear_in_shards; here it is additionally checked to be the row-weighted sum of
    the two 2048-row halves at full size."""
    from odevio_b200.distributed import pose_loss
    B, S, nsub = 4096, 10, 16
    ref, mod = make_pair(cuda_device, bias_std=0.05, ode_solver="dopri5", ode_rtol=1e-3, ode_detach_dt=True)
    ref.train(); mod.train()
    dev = cuda_device
    fv, fi, ts = inputs(B, S, irregular=True, seed=0)
    gts = 0.05 * torch.randn(B, S, 6, generator=torch.Generator().manual_seed(2))

    def run(a, b):
        mod.zero_grad(set_to_none=True)
        fvd = fv[a:b].to(dev).requires_grad_(True)
        fid = fi[a:b].to(dev).requires_grad_(True)
        p, _ = mod(fvd, fid, ts[a:b].to(dev))
        pose_loss(p, gts[a:b].to(dev)).backward()
        assert int(mod.last_status.max().item()) == 0
        return p.detach().cpu(), fvd.grad.cpu(), fid.grad.cpu(), {n: q.grad.clone() for n, q in mod.named_parameters() if q.grad is not None}

    p_all, gfv, gfi, g_all = run(0, B)
    rows = torch.randperm(B, generator=torch.Generator().manual_seed(5))[:nsub]
    fvs = fv[rows].clone().requires_grad_(True)
    fis = fi[rows].clone().requires_grad_(True)
    p_ref, _ = ref(fvs, fis, ts[rows])
    (pose_loss(p_ref, gts[rows]) * (nsub / B)).backward()
    e_pose = rel_err(p_all[rows], p_ref.detach())
    e_fv, e_fi = rel_err(gfv[rows], fvs.grad), rel_err(gfi[rows], fis.grad)
    print(f"configs[3] B=4096: pose {e_pose:.3e}, d loss/d fv {e_fv:.3e}, d loss/d fi {e_fi:.3e} vs oracle autograd on {nsub} rows")
    assert e_pose <= 1e-4 and e_fv <= 2e-4 and e_fi <= 2e-4
    torch.cuda.empty_cache()
    _, _, _, g1 = run(0, B // 2)
    _, _, _, g2 = run(B // 2, B)
    for n in g_all:
        assert rel_err(0.5 * (g1[n] + g2[n]), g_all[n]) <= 2e-5, n


def test_streaming_windows_match_oracle(cuda_device):
    """The streaming driver against the ORACLE (not the kernel against itself): three chained windows with the carried
    state and absolute timestamps, trajectories that restart mid-stream (`reset(rows)`), fp16x3 one-launch kernel."""
    from odevio_b200.streaming import StreamingPoseODERNN
    from helpers import POSE_RTOL, rel_err
    ref, mod = make_pair(cuda_device, bias_std=0.05, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="fp16x3")
    B, W = 12, 4
    fv, fi, ts = inputs(B, 3 * W, irregular=True, seed=6, offset=3.0)
    stream = StreamingPoseODERNN(mod)
    state_ref = None
    for k in range(3):
        a, b = k * W, (k + 1) * W
        if k == 2:                                   # trajectories 0 and 5 restart: zero state, like a fresh sequence
            stream.reset([0, 5])
            state_ref[:, [0, 5]] = 0.0
        with torch.no_grad():
            p_ref, state_ref = ref(fv[:, a:b], fi[:, a:b], ts[:, a:b + 1], prev=state_ref)
        p = stream.step(fv[:, a:b].contiguous().to(cuda_device), fi[:, a:b].contiguous().to(cuda_device),
                        ts[:, a:b + 1].contiguous().to(cuda_device))
        assert rel_err(p.cpu(), p_ref) <= 4 * POSE_RTOL, (k, rel_err(p.cpu(), p_ref))
        assert rel_err(stream.state.cpu(), state_ref) <= 2e-4
