"""Training-loop glue: fused pose loss and flat-bucket clip + Adam against the reference's own
dependencies -- torch.nn.functional.mse_loss, torch.nn.utils.clip_grad_norm_, torch.optim.Adam with the reference's
hyper-parameters (scripts/train_model.py:72-86, src/utils/utils.py:150-157).  PINNED: torch IS what the reference runs."""

import copy

import pytest
import torch

from odevio_b200.distributed import make_optimizer, pose_loss, pose_net_params


def test_flat_views_keep_state_dict_keys_and_values():
    """Host logic (CPU): re-pointing the parameters at one flat buffer keeps names, shapes and values, and later
    load_state_dict() writes through to the buffer (reference checkpoints keep loading)."""
    from oracle.pose_odernn import OraclePoseODERNN, default_opt
    torch.manual_seed(0)
    m = OraclePoseODERNN(default_opt(v_f_len=24, i_f_len=8, ode_hidden_dim=16))
    before = {k: v.clone() for k, v in m.state_dict().items()}
    params = pose_net_params(m)
    flat = torch.zeros(sum(p.numel() for p in params))
    off = 0
    for p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat[off:off + n].view_as(p.data)
        off += n
    after = m.state_dict()
    assert list(before) == list(after) and all(torch.equal(before[k], after[k]) for k in before)
    m.load_state_dict({k: v + 1 for k, v in before.items()})
    assert torch.equal(flat, torch.cat([(before[n] + 1).reshape(-1) for n, _ in
                                        [(n, p) for grp in (True, False) for n, p in m.named_parameters()
                                         if n.startswith("regressor") == grp]]))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [(4, 3), (1024, 10), (333, 7)])
def test_fused_pose_loss_matches_torch(cuda_device, n):
    from odevio_b200.training import fused_pose_loss
    g = torch.Generator().manual_seed(1)
    poses = torch.randn(*n, 6, generator=g).to(cuda_device).requires_grad_(True)
    gts = (0.1 * torch.randn(*n, 6, generator=g)).to(cuda_device)
    loss, parts = fused_pose_loss(poses, gts, with_parts=True)
    (3.0 * loss).backward()
    p2 = poses.detach().clone().requires_grad_(True)
    want = pose_loss(p2, gts)
    (3.0 * want).backward()
    assert abs(loss.item() - want.item()) <= 2e-6 * abs(want.item())
    assert abs(parts[1].item() - torch.nn.functional.mse_loss(p2[..., :3], gts[..., :3]).item()) <= 2e-6 * parts[1].item()
    assert ((poses.grad - p2.grad).abs().max() / p2.grad.abs().max()).item() <= 2e-6


@pytest.mark.gpu
def test_fused_adam_matches_torch_adam_with_clip(cuda_device):
    """Five optimisation steps of the flat-bucket clip + Adam vs clip_grad_norm_ + torch.optim.Adam on identical
    parameters and gradients; one step with a gradient norm above the clip threshold, the others below."""
    import odevio_b200
    from odevio_b200.training import FusedPoseNetAdam
    from oracle.pose_odernn import default_opt
    opt = default_opt()
    torch.manual_seed(0)
    a = odevio_b200.PoseODERNN(opt).to(cuda_device)
    b = copy.deepcopy(a)
    fused = FusedPoseNetAdam(a, lr=1e-3, weight_decay=5e-5, max_norm=5.0)
    ref_opt = make_optimizer(b, lr=1e-3, weight_decay=5e-5)
    keys = list(a.state_dict())
    g = torch.Generator().manual_seed(3)
    assert [len(g["params"]) for g in fused.param_groups] == [len(g["params"]) for g in ref_opt.param_groups]
    for step in range(5):
        if step == 3:
            # the reference's epoch loop re-schedules ONLY group 0 (scripts/train_model.py:215-216; group 0 = the
            # non-regressor parameters, utils/utils.py:116-119): the regressor keeps the warm-up rate
            fused.param_groups[0]["lr"] = ref_opt.param_groups[0]["lr"] = 1e-5
        scale = 50.0 if step == 2 else 0.01                       # step 2 exceeds max_norm = 5
        for pa, pb in zip(pose_net_params(a), pose_net_params(b)):
            gr = (scale * torch.randn(pa.shape, generator=g) / pa.numel() ** 0.5).to(cuda_device)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        fused.gather_grads()
        fused.step()
        norm = torch.nn.utils.clip_grad_norm_(pose_net_params(b), max_norm=5.0)
        ref_opt.step()
        assert abs(fused.norm_coef[0].item() - norm.item()) <= 1e-5 * norm.item()
        assert (fused.norm_coef[1].item() < 1.0) == (norm.item() > 5.0)
        for pa, pb in zip(pose_net_params(a), pose_net_params(b)):
            assert ((pa - pb).abs().max() / pb.abs().max().clamp_min(1e-12)).item() <= 1e-5, step
    assert list(a.state_dict()) == keys                            # checkpoint keys unchanged by the flat views
    # resume: moments / step / group rates survive a state_dict round trip into a fresh optimiser on the same model
    sd = fused.state_dict()
    again = FusedPoseNetAdam(a)
    again.load_state_dict(sd)
    assert again.step_count == 5 and again.param_groups[0]["lr"] == 1e-5 and again.param_groups[1]["lr"] == 1e-3
    assert torch.equal(again.exp_avg, fused.exp_avg)
    # a parameter re-assigned behind the optimiser's back is detected instead of silently ignored
    a.regressor[2].bias.data = a.regressor[2].bias.data.clone()
    with pytest.raises(Exception, match="no longer aliases"):
        again.step()


@pytest.mark.gpu
def test_fused_train_step_matches_reference_glue(cuda_device):
    """Whole step: fused forward/backward + fused glue vs fused forward/backward + the torch glue of
    odevio_b200.distributed.train_step (mse_loss, clip_grad_norm_, torch.optim.Adam): same loss, same parameters."""
    import odevio_b200
    from helpers import inputs
    from odevio_b200.distributed import train_step
    from odevio_b200.training import FusedPoseNetAdam, fused_train_step
    from oracle.pose_odernn import default_opt
    opt = default_opt(ode_solver="rk4")
    torch.manual_seed(0)
    a = odevio_b200.PoseODERNN(opt).to(cuda_device).train()
    b = copy.deepcopy(a)
    fused = FusedPoseNetAdam(a)
    ref_opt = make_optimizer(b)
    fv, fi, ts = (t.to(cuda_device) for t in inputs(16, S=4))
    gts = (0.1 * torch.randn(16, 4, 6, generator=torch.Generator().manual_seed(2))).to(cuda_device)
    p0 = [p.detach().clone() for p in pose_net_params(b)]
    steps, lr = 3, 1e-3
    for _ in range(steps):
        la = fused_train_step(a, fused, fv, fi, ts, gts)
        lb = train_step(b, ref_opt, fv, fi, ts, gts)
        assert abs(la.item() - lb.item()) <= 1e-5 * abs(lb.item())
    # Adam normalises every element's update to ~lr * sign(g): elements whose gradient is at the fp32 noise level of the
    # backward (a cancellation residue ~1e-6 of the tensor's largest gradient; the two loss backwards round differently)
    # legitimately move differently.  So: the update VECTORS agree in the L2 sense, no element is off by more than the
    # steps could move it, and the identical-gradient test above pins the optimiser arithmetic itself to 1e-5.
    da = torch.cat([(x.detach() - z).reshape(-1) for x, z in zip(pose_net_params(a), p0)])
    db = torch.cat([(y.detach() - z).reshape(-1) for y, z in zip(pose_net_params(b), p0)])
    assert ((da - db).norm() / db.norm()).item() <= 2e-2, ((da - db).norm() / db.norm()).item()
    assert (da - db).abs().max().item() <= 2 * lr * steps


@pytest.mark.gpu
def test_peer_allreduce_adam_kernel_world1(cuda_device):
    """odevio_allreduce_adam_peer with world = 1 (its own peer): the one-launch reduce + clip + Adam must equal the
    three-launch odevio_adam_step_groups on identical gradients, incl. a clipped step and the two group rates."""
    import odevio_b200
    from odevio_b200.training import FusedPoseNetAdam, PeerFusedPoseNetAdam
    from oracle.pose_odernn import default_opt
    torch.manual_seed(0)
    a = odevio_b200.PoseODERNN(default_opt()).to(cuda_device)
    b = copy.deepcopy(a)
    peer = PeerFusedPoseNetAdam(a, lr=1e-3)
    base = FusedPoseNetAdam(b, lr=1e-3)
    peer.param_groups[0]["lr"] = base.param_groups[0]["lr"] = 1e-4
    g = torch.Generator().manual_seed(3)
    for step in range(4):
        scale = 50.0 if step == 2 else 0.01
        for pa, pb in zip(pose_net_params(a), pose_net_params(b)):
            gr = (scale * torch.randn(pa.shape, generator=g) / pa.numel() ** 0.5).to(cuda_device)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        peer.gather_grads(); peer.step_allreduce()
        base.gather_grads(); base.step()
        assert abs(peer.norm_coef[0].item() - base.norm_coef[0].item()) <= 1e-6 * base.norm_coef[0].item()
        for pa, pb in zip(pose_net_params(a), pose_net_params(b)):
            assert ((pa - pb).abs().max() / pb.abs().max().clamp_min(1e-12)).item() <= 1e-6, step
    assert torch.allclose(peer.exp_avg, base.exp_avg, rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_peer_allreduce_adam_two_gpus():
    """Two ranks over NVLink peer memory against NCCL all-reduce + the three-launch step (tools/peer_step_check.py);
    skipped on a single-GPU box."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(root, "tools", "peer_step_check.py")],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, PYTHONPATH=root))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PEER_STEP_OK" in r.stdout
