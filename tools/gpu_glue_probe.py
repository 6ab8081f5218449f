import copy, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import odevio_b200
from helpers import inputs
from odevio_b200.distributed import make_optimizer, pose_loss, pose_net_params, train_step
from odevio_b200.training import FusedPoseNetAdam, fused_pose_loss, fused_train_step
from oracle.pose_odernn import default_opt
dev = torch.device("cuda:0")
opt = default_opt(ode_solver="rk4")
torch.manual_seed(0)
a = odevio_b200.PoseODERNN(opt).to(dev).train()
b = copy.deepcopy(a)
fused = FusedPoseNetAdam(a)
ref_opt = make_optimizer(b)
fv, fi, ts = (t.to(dev) for t in inputs(16, S=4))
gts = (0.1 * torch.randn(16, 4, 6, generator=torch.Generator().manual_seed(2))).to(dev)
def rel(x, y): return ((x - y).abs().max() / y.abs().max().clamp_min(1e-30)).item()
for step in range(3):
    fused.zero_grad(); ref_opt.zero_grad(set_to_none=True)
    pa, _ = a(fv, fi, ts); pb, _ = b(fv, fi, ts)
    la = fused_pose_loss(pa, gts); lb = pose_loss(pb, gts)
    la.backward(); lb.backward()
    print(f"step {step}: pose rel {rel(pa, pb):.2e} loss {la.item():.6f} {lb.item():.6f}")
    worst = max(((rel(x.grad, y.grad), n) for (n, x), (_, y) in zip(a.named_parameters(), b.named_parameters()) if y.grad is not None), default=None)
    print("   worst grad rel diff", worst)
    fused.gather_grads(); fused.step()
    norm = torch.nn.utils.clip_grad_norm_(pose_net_params(b), max_norm=5.0); ref_opt.step()
    print("   norm", fused.norm_coef.tolist(), norm.item())
    worst = max((rel(x, y), n) for (n, x), (_, y) in zip(a.named_parameters(), b.named_parameters()))
    print("   worst param rel diff", worst)
    # element-level: where is the biggest param diff and what was its gradient
    n = worst[1]; x = dict(a.named_parameters())[n]; y = dict(b.named_parameters())[n]
    idx = (x - y).abs().argmax(); print("   at", n, idx.item(), "pa", x.reshape(-1)[idx].item(), "pb", y.reshape(-1)[idx].item(), "grad", y.grad.reshape(-1)[idx].item())
