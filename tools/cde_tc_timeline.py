"""Where the time of the tensor-core CDE forward goes: clock sums of CTA 0 over all evaluations of one configs[2] launch
(cde_tc.cu: g_cde_tc_dbg).  Slots: 0 evaluations; row phase, cumulative from the start of an evaluation: 1 stage argument,
2 + dX/dt, 3 + Hc x Hc Linears, 4 + image / K_out stores (end of row phase), 5 + first grid barrier; feature phase, from its
start: 6 epilogue warps done, 16 producer done, 24 MMA issuer done, 7 + second grid barrier; 8 epilogue wait for accumulators,
25 / 26 MMA issuer waits (operands / accumulator buffer); 15 whole kernel."""
import ctypes as C
import sys
from types import SimpleNamespace

import torch

import odevio_b200
from odevio_b200 import _lib, synth

B, S, Hc = 1024, 10, 128
interp = sys.argv[1] if len(sys.argv) > 1 else "cubic"
opt = SimpleNamespace(v_f_len=Hc // 2, i_f_len=Hc // 2, fuse_method="cat", cde_hidden_dim=Hc, cde_fn_num_layers=3,
                      cde_num_layers=3, cde_activation_fn="tanh", cde_solver="dopri5", adjoint=False, cde_interp=interp,
                      cde_precision="fp16x3")
model = odevio_b200.PoseCDE(opt)
torch.manual_seed(0)
for m in model.modules():
    if isinstance(m, torch.nn.Linear):
        torch.nn.init.kaiming_normal_(m.weight.data); m.bias.data.zero_()
model = model.cuda().train()
fv, fi = synth.features(B, S, Hc // 2, Hc // 2, seed=0)
ts = synth.timestamps(B, S, irregular=True, seed=0)
fv, fi, ts = (0.2 * fv).cuda(), (0.2 * fi).cuda(), ts.cuda()
with torch.no_grad():
    for _ in range(3):
        model(fv, fi, ts)
torch.cuda.synchronize()
lib = _lib.load()
lib.odevio_debug_cde_tc_timeline.restype = C.c_int32
buf = (C.c_longlong * 32)()
assert lib.odevio_debug_cde_tc_timeline(buf) == 0
d = list(buf)
n = max(d[0], 1)
print("stats", model.last_stats.tolist(), "evaluations", d[0], "kernel clk", d[15])
names = {1: "row: stage argument", 2: "row: + dX/dt", 3: "row: + Linears", 4: "row: + stores (end)", 5: "row: + grid barrier",
         6: "feat: epilogue done", 16: "feat: producer done", 24: "feat: MMA issuer done", 7: "feat: + grid barrier",
         9: "row: wait for Linear weights", 10: "row: Linear compute (sum)", 11: "row: Linear + named barrier (sum)",
         8: "epilogue wait tfull", 25: "MMA wait operands", 26: "MMA wait accumulator buffer"}
for k in (1, 2, 3, 4, 5, 9, 10, 11, 6, 16, 24, 7, 8, 25, 26):
    print(f"  {names[k]:32s} {d[k] / n:10.0f} clk per evaluation")
