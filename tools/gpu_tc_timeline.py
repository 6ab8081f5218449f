"""Development: clock64 timeline of the tensor-core solver kernel (build variant -DODEVIO_FT_TIMELINE): the last
solver iteration of cluster 0 / CTA 0 on one dopri5 interval."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["ODEVIO_LIB_PATH"] = os.path.join(ROOT, "odevio_b200", "lib", "libodevio_b200.timeline.so")
import torch
from helpers import make_pair
from odevio_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.load()
ref, mod = make_pair(dev, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="tf32x3", bias_std=0.05)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
g = torch.Generator().manual_seed(1)
y = (0.3 * torch.randn(M, 768, generator=g)).to(dev)
ts = torch.stack([torch.zeros(M), torch.full((M,), 0.1)], 1).to(dev)
with torch.no_grad():
    for _ in range(2): mod.evolve_state(y, ts)
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
lib.odevio_debug_tc_timeline.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.odevio_debug_tc_timeline(buf), "steps", mod.last_stats[0, 0, :, 0].max().item())
t = list(buf)
print("iteration begin -> last stage begin:", t[41] - t[40], "(5 earlier stages)")
print("stage_input:", t[42] - t[41], " fence.proxy.async:", t[43] - t[42], " cluster barrier:", t[44] - t[43])
prev = t[44]
for l in range(3):
    b = 8 + l * 8
    names = ["first chunk", "mma issued", "accum ready", "epilogue done", "barrier passed"]
    print(f"layer {l}: " + "  ".join(f"{n} +{t[b+i]-prev}" for i, n in enumerate(names)))
    prev = t[b + 4]
print("after layer returns (rel. to barrier passed):", [t[52 + l] - t[8 + l * 8 + 4] for l in range(3)], " loop exit:", t[45] - t[54])
print("stage total:", t[45] - t[41])
print("error pass:", t[46] - t[45], " bar+partial store:", t[47] - t[46], " cluster barrier:", t[48] - t[47],
      " controller + syncthreads:", t[49] - t[48], " commit:", t[50] - t[49])
print("whole iteration:", t[50] - t[40])
