"""Development: clock64 timeline of the 3xFP16 solver kernel (build variant -DODEVIO_H3_TIMELINE): the last solver
iteration of cluster 0 / CTA 0 on one dopri5 interval (last stage of the iteration)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["ODEVIO_LIB_PATH"] = os.path.join(ROOT, "odevio_b200", "lib", "libodevio_b200.%s.so" % (sys.argv[2] if len(sys.argv) > 2 else "h3timeline"))
import torch
from helpers import make_pair
from odevio_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.load()
ref, mod = make_pair(dev, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="fp16x3", bias_std=0.05)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
g = torch.Generator().manual_seed(1)
y = (0.3 * torch.randn(M, 768, generator=g)).to(dev)
ts = torch.stack([torch.zeros(M), torch.full((M,), 0.1)], 1).to(dev)
with torch.no_grad():
    for _ in range(2): mod.evolve_state(y, ts)
torch.cuda.synchronize()
buf = (C.c_longlong * 96)()
lib.odevio_debug_h3_timeline.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.odevio_debug_h3_timeline(buf), "steps", mod.last_stats[0, 0, :, 0].max().item())
t = list(buf)
print("iteration begin -> last stage begin:", t[1] - t[0], "(5 earlier stages)")
print("stage_input:", t[2] - t[1], " cluster barrier:", t[3] - t[2])
prev = t[3]
for l in range(4):
    b = 16 + 10 * l
    print(f"layer {l}: first chunk +{t[b]-prev}  last chunk landed +{t[b+5]-prev}  mma issued +{t[b+1]-prev}  accum ready +{t[b+2]-prev}  "
          f"epilogue done +{t[b+3]-prev}  barrier passed +{t[b+4]-prev}   | starved {t[b+6]} (after chunk 0)")
    prev = t[b + 4]
print("stage total:", t[4] - t[1])
print("error pass:", t[5] - t[4], " bar+partial store:", t[6] - t[5], " cluster barrier:", t[7] - t[6],
      " controller + syncthreads:", t[8] - t[7], " commit:", t[9] - t[8])
print("whole iteration:", t[9] - t[0])
