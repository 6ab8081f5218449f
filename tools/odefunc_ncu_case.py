"""Small driver for ncu: a few tensor-core ODEFunc.forward evaluations at M = 32768 rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, odevio_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
f = odevio_b200.ODEFunc(768, 512, 3, "tanh").to(dev)
x = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 768, device=dev)
with torch.no_grad():
    for _ in range(4):
        f(None, x)
torch.cuda.synchronize()
print("ok")
