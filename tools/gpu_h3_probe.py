"""Development probe for the 3xFP16 solver kernel (odernn_h3.cu): evolve_state vs the FFMA kernel, error maps, timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import odevio_b200
from oracle.pose_odernn import default_opt
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_pair, inputs, rel_err

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "parity"


def pair(prec, **over):
    ref, mod = make_pair(dev, ode_precision=prec, bias_std=0.05, **over)
    return ref, mod


if mode in ("parity", "all"):
    for solver, sub in (("euler_fixed", 1), ("rk4", 1), ("dopri5", 1)):
        over = dict(ode_solver="rk4" if solver != "dopri5" else "dopri5", ode_substeps=sub)
        if solver == "dopri5":
            over["ode_rtol"] = 1e-3
        ref, mod = pair("fp16x3", **over)
        _, mod_f = pair("fp32", **over)
        g = torch.Generator().manual_seed(11)
        for B in (16, 64, 100):
            state = (0.5 * torch.randn(B, mod.f_len, generator=g)).to(dev)
            t0 = torch.rand(B, generator=g)
            ts = torch.stack([t0, t0 + 0.1 + 0.3 * torch.rand(B, generator=g)], 1).to(dev)
            with torch.no_grad():
                got = mod.evolve_state(state, ts)
                torch.cuda.synchronize()
                want = mod_f.evolve_state(state, ts)
                torch.cuda.synchronize()
            err = (got - want).abs()
            print(f"{solver} B={B}: rel_err {rel_err(got.cpu(), want.cpu()):.3e}  status {int(mod.last_status.max())} "
                  f"steps h3 {mod.last_stats[0,0,:4,0].tolist()} fma {mod_f.last_stats[0,0,:4,0].tolist()}", flush=True)
            if rel_err(got.cpu(), want.cpu()) > 1e-4:
                e = err.cpu()
                print("  per-row max err (first 16 rows):", [f"{v:.1e}" for v in e.max(1).values[:16].tolist()])
                fe = e.max(0).values
                print("  per-feature-block(32) max err:", [f"{v:.1e}" for v in fe.view(-1, 32).max(1).values.tolist()])
                d = (want - state).cpu(); dg = (got - state).cpu()
                print("  |want-y0| max", d.abs().max().item(), "|got-y0| max", dg.abs().max().item())
                print("  sample want-y0", d[0, :8].tolist()); print("  sample got-y0 ", dg[0, :8].tolist())
                break

if mode in ("time", "all"):
    for prec in ("fp16x3", "tf32x3"):
        ref, mod = pair(prec, ode_solver="dopri5", ode_rtol=1e-3)
        fv, fi, ts = inputs(1024, 10, irregular=True, seed=0)
        fv, fi, ts = fv.to(dev), fi.to(dev), ts.to(dev)
        with torch.no_grad():
            for _ in range(3):
                p, h = mod(fv, fi, ts)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                p, h = mod(fv, fi, ts)
            e1.record(); torch.cuda.synchronize()
        st = mod.last_stats.float()
        print(f"{prec}: {e0.elapsed_time(e1)/5:.2f} ms per forward (B=1024), mean steps {st[...,0].mean():.2f}, status {int(mod.last_status.max())}", flush=True)
        if prec == "fp16x3":
            p16 = p.clone()
        else:
            print("  fp16x3 vs tf32x3 poses rel", rel_err(p16.cpu(), p.cpu()))
    # fixed-step evaluation cost: rk4 with many substeps on 2048 rows
    for prec in ("fp16x3", "tf32x3"):
        ref, mod = pair(prec, ode_solver="rk4", ode_substeps=16)
        g = torch.Generator().manual_seed(4)
        state = (0.5 * torch.randn(2048, mod.f_len, generator=g)).to(dev)
        t0 = torch.rand(2048, generator=g)
        ts = torch.stack([t0, t0 + 0.3], 1).to(dev)
        with torch.no_grad():
            mod.evolve_state(state, ts); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mod.evolve_state(state, ts); e1.record(); torch.cuda.synchronize()
        print(f"{prec}: rk4 x16 substeps on 2048 rows: {e0.elapsed_time(e1):.3f} ms = {e0.elapsed_time(e1)*1e3/64:.1f} us per ODEFunc evaluation", flush=True)
