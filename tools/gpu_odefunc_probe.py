"""Developer probe: tensor-core ODEFunc.forward timing (CUDA events) next to torch's fp32 / tf32 MLP."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import odevio_b200
dev = torch.device("cuda:0")
D, H, n = 768, 512, 3
torch.manual_seed(0)
f = odevio_b200.ODEFunc(D, H, n, "tanh")
for m in f.net:
    if isinstance(m, torch.nn.Linear):
        torch.nn.init.kaiming_normal_(m.weight.data); m.bias.data.normal_(0, 0.05)
f = f.to(dev)
flops_row = 2 * (D * H + (n - 1) * H * H + H * D)
for M in (2048, 8192, 32768, 131072):
    x = torch.randn(M, D, device=dev)
    with torch.no_grad():
        for name, fn in (("tcgen05 3xTF32", lambda: f(None, x)), ("torch fp32", lambda: f.net(x))):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20 if M <= 8192 else 5
            e0.record()
            for _ in range(iters): out = fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"M={M:7d} {name:16s} {ms*1e3:9.1f} us/eval  {M*flops_row/ms/1e9:8.2f} TFLOP/s (algorithmic)", flush=True)
        ref = f.net(x.double().cpu()[:256].to(dev).double()) if False else None
