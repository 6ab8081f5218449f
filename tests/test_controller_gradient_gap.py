"""How far the kernels' gradient (discretise-then-optimise with the accepted step sizes as CONSTANTS) is from the
reference's (torchode's AutoDiffAdjoint back-propagates through the step-size controller too), measured on the ORACLE --
no GPU involved: oracle with ``detach_dt`` (the kernels' exact target, tests/test_odernn_backward_gpu.py) against the oracle
differentiating through its controller, at rtol = 1e-3 and at the reference's rtol = 1e-2 (PoseODERNN.py:57).

Finding (printed, recorded in DESIGN.md 4.2): at rtol = 1e-3 the two gradients agree to ~2e-5 (max-norm relative) on the
3-interval case below; at the reference's rtol = 1e-2 with dt0 = 1e-4 the first error ratios underflow to exactly 0 and the
through-controller gradient of the restated controller is NOT FINITE (d/dx x^(-1/5) at 0 times the clamp's zero) -- i.e. at
its own tolerances the reference's extra term is either negligible or undefined in fp32, so the constant-step gradient is
the well-defined object to match.  [Depends on the recalled controller arithmetic, oracle/torchode_like.py.]"""

import math

import torch

from helpers import inputs, rel_err
from oracle.modules import deepvio_initialization
from oracle.pose_odernn import OraclePoseODERNN, default_opt


def _loss_grads(model, fv, fi, ts, gts):
    model.zero_grad(set_to_none=True)
    pose, _ = model(fv, fi, ts)
    loss = 100 * torch.nn.functional.mse_loss(pose[:, :, :3], gts[:, :, :3]) + \
        torch.nn.functional.mse_loss(pose[:, :, 3:], gts[:, :, 3:])
    loss.backward()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def _gap(rtol):
    opt = default_opt(ode_rtol=rtol, ode_detach_dt=True)
    torch.manual_seed(0)
    ref = OraclePoseODERNN(opt)
    deepvio_initialization(ref)
    ref.train()
    g = torch.Generator().manual_seed(7)
    for n, p in ref.named_parameters():
        if n.endswith("bias"):
            p.data.normal_(0.0, 0.05, generator=g)
    fv, fi, ts = inputs(8, 3, irregular=True, seed=2)
    gts = 0.1 * torch.randn(8, 3, 6, generator=torch.Generator().manual_seed(9))
    ref.ctrl.detach_dt = False
    g_full = _loss_grads(ref, fv, fi, ts, gts)
    ref.ctrl.detach_dt = True
    g_const = _loss_grads(ref, fv, fi, ts, gts)
    assert all(torch.isfinite(v).all() for v in g_const.values())
    return max(rel_err(g_const[k], g_full[k]) for k in g_full)


def test_controller_gradient_gap_is_reported():
    gap3, gap2 = _gap(1e-3), _gap(1e-2)
    print(f"controller-gradient gap, constant-dt vs through-controller (oracle): rtol 1e-3: {gap3:.3e}, rtol 1e-2: {gap2:.3e}")
    assert gap3 <= 5e-2
    assert (not math.isfinite(gap2)) or gap2 <= 0.5
