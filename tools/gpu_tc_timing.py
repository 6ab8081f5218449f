"""Timing probe of the tensor-core solver kernel alone: PoseODERNN.evolve_state (L = 1, rows = B) with fixed-step
rk4 (4 * substeps evaluations per call) for several row counts; prints us per vector-field evaluation."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_pair
from odevio_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load()
sub = int(os.environ.get("SUB", "8"))
for prec in ("tf32x3", "fp32"):
    ref, mod = make_pair(dev, ode_solver="rk4", ode_substeps=sub, ode_precision=prec, bias_std=0.05)
    for M in (128, 512, 1024, 1920, 2048, 4096):
        g = torch.Generator().manual_seed(1)
        y = (0.3 * torch.randn(M, 768, generator=g)).to(dev)
        ts = torch.stack([torch.zeros(M), torch.full((M,), 0.1)], 1).to(dev)
        with torch.no_grad():
            mod.evolve_state(y, ts)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): mod.evolve_state(y, ts)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        geo = (C.c_int32 * 3)()
        lib.odevio_debug_tc_geometry(geo)
        print(f"{prec} rows={M}: {ms:.3f} ms per call, {ms * 1e3 / (4 * sub):.1f} us per evaluation "
              f"({M * 2.62144e6 * 4 * sub / ms / 1e9:.1f} TFLOP/s), clusters {geo[0]} max {geo[1]}", flush=True)
