"""Tiny end-to-end case for compute-sanitizer: forward (rk4 + dopri5), fused backward, CDE forward,
tensor-core ODEFunc.forward -- every kernel of the library once, smallest shapes that reach them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import odevio_b200
from odevio_b200 import synth
from types import SimpleNamespace
dev = torch.device("cuda:0")
opt = SimpleNamespace(v_f_len=512, i_f_len=256, fuse_method="cat", ode_hidden_dim=512, ode_fn_num_layers=3,
                      ode_activation_fn="tanh", ode_solver="dopri5", ode_rnn_type="rnn", rnn_num_layers=2,
                      rnn_hidden_dim=1024, rnn_dropout_out=0.0)
torch.manual_seed(0)
m = odevio_b200.PoseODERNN(opt).to(dev)
fv, fi = synth.features(9, 2)
ts = synth.timestamps(9, 2, irregular=True)
fv, fi, ts = fv.to(dev), fi.to(dev), ts.to(dev)
with torch.no_grad():
    m.eval(); m(fv, fi, ts)
m.train()
p, h = m(fv, fi, ts)
(p ** 2).mean().backward()
opt.ode_rnn_type = "gru"; opt.ode_solver = "rk4"
g = odevio_b200.PoseODERNN(opt).to(dev).train()
p, h = g(fv, fi, ts)
(p ** 2).mean().backward()
copt = SimpleNamespace(v_f_len=16, i_f_len=16, fuse_method="cat", cde_hidden_dim=32, cde_fn_num_layers=2, cde_num_layers=3,
                       cde_activation_fn="tanh", cde_solver="dopri5", adjoint=False, cde_interp="cubic", cde_rtol=1e-2)
c = odevio_b200.PoseCDE(copt).to(dev).train()
cf, ci = synth.features(9, 3, 16, 16)
with torch.no_grad():
    c(0.2 * cf.to(dev), 0.2 * ci.to(dev), synth.timestamps(9, 3).to(dev))
    f = odevio_b200.ODEFunc(768, 512, 3, "tanh").to(dev)
    f(None, torch.randn(130, 768, device=dev))
    f(None, torch.randn(128 * 20, 768, device=dev))        # NC = 4 instantiation
    odevio_b200.CDEFunc(9, 8, 2, "tanh").to(dev)(None, torch.randn(5, 8, device=dev))
torch.cuda.synchronize()
print("sanitizer case ok")
