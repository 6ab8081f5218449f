"""Development: per-phase clock64 timeline of odefunc_tc_kernel (build variant with -DODEVIO_FT_TIMELINE)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ODEVIO_LIB_PATH"] = os.path.join(ROOT, "odevio_b200", "lib", "libodevio_b200.%s.so" % os.environ.get("FT_VARIANT", "timeline"))
import torch, odevio_b200
from odevio_b200 import _lib
dev = torch.device("cuda:0")
f = odevio_b200.ODEFunc(768, 512, 3, "tanh").to(dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = torch.randn(M, 768, device=dev)
with torch.no_grad():
    for _ in range(3): f(None, x)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * 64)()
lib.odevio_debug_odefunc_timeline.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.odevio_debug_odefunc_timeline(buf))
t = list(buf); t0 = t[0]
print(f"convert done -> barrier: {t[1]-t[0]} clk")
prev = t[1]
for l in range(4):
    b = 8 + l * 8
    names = ["first chunk", "mma issued", "accum ready", "epilogue done", "barrier passed"]
    print(f"layer {l}: " + "  ".join(f"{n} +{t[b+i]-prev}" for i, n in enumerate(names)))
    prev = t[b + 4]
print("total clk", t[8 + 3 * 8 + 4] - t[0])
