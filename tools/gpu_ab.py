"""A/B timing of library variants (developer tool): python tools/gpu_ab.py v1 v2 ..."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import torch
from helpers import make_pair, inputs
dev = torch.device("cuda:0")
def timing(name, B, S=10, iters=3, **over):
    ref, mod = make_pair(dev, bias_std=0.05, **over)
    fv, fi, ts = [t.to(dev) for t in inputs(B, S, irregular=True)]
    with torch.no_grad():
        mod(fv, fi, ts); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): mod(fv, fi, ts)
        e1.record(); torch.cuda.synchronize()
    print("   %%-34s %%8.2f ms" %% (name, e0.elapsed_time(e1) / iters), flush=True)
timing("dopri5 rtol1e-3 B=1024 rt8", 1024, ode_rtol=1e-3, ode_rows_per_tile=8)
timing("rk4 B=1024 rt8", 1024, ode_solver="rk4", ode_rows_per_tile=8)
timing("rk4 B=2048 rt16", 2048, ode_solver="rk4", ode_rows_per_tile=16)
''' % (root, root)
for v in sys.argv[1:]:
    env = dict(os.environ)
    if v != "default":
        env["ODEVIO_LIB_PATH"] = os.path.join(root, "odevio_b200", "lib", f"libodevio_b200.{v}.so")
    print(f"== variant {v}", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env, timeout=300)
