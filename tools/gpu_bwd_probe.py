"""Developer probe: backward parity per gradient tensor + timing.  python tools/gpu_bwd_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import inputs, make_pair, rel_err
from test_odernn_backward_gpu import _grads

dev = torch.device("cuda:0")
def case(name, B, S, **over):
    ref, mod = make_pair(dev, bias_std=0.05, ode_detach_dt=True, **over)
    ref.train(); mod.train()
    fv, fi, ts = inputs(B, S, irregular=True, seed=2)
    g = torch.Generator().manual_seed(9)
    gts = 0.1 * torch.randn(B, S, 6, generator=g)
    l_ref, g_ref, p_ref = _grads(ref, fv, fi, ts, gts, None, 0.0)
    l_gpu, g_gpu, p_gpu = _grads(mod, fv.to(dev), fi.to(dev), ts.to(dev), gts.to(dev), None, 0.0)
    print(f"[{name}] loss ref {l_ref:.6f} gpu {l_gpu:.6f} pose_err {rel_err(p_gpu, p_ref):.2e} status {int(mod.last_status.max())}")
    st = mod.last_stats.cpu().long()
    neq = (st[..., 0] != ref.last_stats["n_steps"]) | (st[..., 1] != ref.last_stats["n_accepted"])
    print(f"   step-count mismatches {int(neq.sum())}/{neq.numel()}  mean steps {st[...,0].float().mean():.2f}")
    for k in g_ref:
        print(f"   {k:28s} rel_err {rel_err(g_gpu[k], g_ref[k]):.3e}  |ref| {g_ref[k].abs().max():.3e}")

def timing(B, iters=2, **over):
    ref, mod = make_pair(dev, bias_std=0.05, **over)
    mod.train()
    fv, fi, ts = inputs(B, 10, irregular=True)
    fv, fi, ts = fv.to(dev), fi.to(dev), ts.to(dev)
    gts = torch.zeros(B, 10, 6, device=dev)
    for k in range(iters + 1):
        mod.zero_grad(set_to_none=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        pose, h = mod(fv, fi, ts)
        e[1].record()
        loss = ((pose - gts) ** 2).mean()
        loss.backward()
        e[2].record(); torch.cuda.synchronize()
        print(f"[timing B={B}] fwd {e[0].elapsed_time(e[1]):.1f} ms  bwd {e[1].elapsed_time(e[2]):.1f} ms  wall {(time.perf_counter()-t0)*1e3:.1f} ms  "
              f"mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB", flush=True)

cases = sys.argv[1:] or ["parity"]
if "parity" in cases:
    case("rk4 B=8 S=2", 8, 2, ode_solver="rk4")
    case("dopri5 B=8 S=3", 8, 3, ode_solver="dopri5", ode_rtol=1e-3)
if "relu" in cases:
    case("relu rk4 B=6 S=3", 6, 3, ode_solver="rk4", ode_activation_fn="relu")
    case("relu dopri5 B=6 S=3", 6, 3, ode_activation_fn="relu")
    case("leaky dopri5 B=6 S=3", 6, 3, ode_activation_fn="leaky_relu")
if "timing" in cases:
    timing(int(os.environ.get("PROBE_B", "1024")), iters=3, ode_rtol=1e-3)
