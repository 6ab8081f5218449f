"""Developer probe for the CDE kernel.  python tools/gpu_cde_probe.py [parity|timing]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_cde_gpu import make_pair, run, data
dev = torch.device("cuda:0")
cases = sys.argv[1:] or ["parity"]
if "parity" in cases:
    for name, kw, dkw in [
        ("linear dopri5 regular", dict(cde_fn_num_layers=2), dict(irregular=False)),
        ("linear dopri5 irregular", dict(cde_fn_num_layers=2), dict(irregular=True)),
        ("cubic dopri5", dict(cde_fn_num_layers=2, cde_interp="cubic"), dict(irregular=True)),
        ("linear rk4", dict(cde_fn_num_layers=2, cde_solver="rk4"), dict(irregular=True)),
        ("cubic rk4 step .25", dict(cde_fn_num_layers=2, cde_solver="rk4", cde_interp="cubic", cde_step_size=0.25), dict(irregular=True)),
    ]:
        ref, mod = make_pair(dev, **kw)
        out = run(ref, mod, *data(12, 10, 32, **dkw), dev)
        print(f"[{name}] pose_err {out['pose_err']:.3e} z0_err {out['z0_err']:.3e} stats {out['stats']} ref {out['ref_stats']}", flush=True)
if "timing" in cases:
    for interp in ("linear", "cubic"):
        ref, mod = make_pair(dev, Hc=128, cde_fn_num_layers=3, cde_interp=interp, train=True)
        fv, fi, ts = data(1024, 10, 128, True)
        fv, fi, ts = fv.to(dev), fi.to(dev), ts.to(dev)
        with torch.no_grad():
            for k in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); mod(fv, fi, ts); e1.record(); torch.cuda.synchronize()
                print(f"[timing {interp} B=1024 Hc=128] {e0.elapsed_time(e1):.2f} ms stats {mod.last_stats.tolist()}", flush=True)
