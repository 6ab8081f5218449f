"""Development: run the configs[1] forward (one odernn_h3_kernel launch) N times and check that poses, final states and step
statistics are bit-identical run to run (a cross-CTA ordering bug in the cluster kernel shows up as run-to-run differences)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import odevio_b200
from odevio_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
w = bench.WORKLOAD
model = odevio_b200.PoseODERNN(bench.make_opt())
bench.init_like_deepvio(model, seed=0)
model = model.to(dev).eval()
fv, fi = synth.features(w["B"], w["S"], w["v_f_len"], w["i_f_len"], seed=0)
ts = synth.timestamps(w["B"], w["S"], irregular=True, seed=0)
fv, fi, ts = fv.to(dev), fi.to(dev), ts.to(dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ref = None
bad = 0
with torch.no_grad():
    for k in range(N):
        if k % 2:
            flush.fill_(k & 255)          # vary the L2 state between runs
        p, h = model(fv, fi, ts)
        st = model.last_stats.clone()
        if ref is None:
            ref = (p.clone(), h.clone(), st)
        elif not (torch.equal(p, ref[0]) and torch.equal(h, ref[1]) and torch.equal(st, ref[2])):
            bad += 1
torch.cuda.synchronize()
print(f"{N} forwards of configs[1]: {bad} differ from the first (status max {int(model.last_status.max().item())})")
sys.exit(1 if bad else 0)
