"""Probe: tensor-core solver path vs the FMA kernel vs the oracle (step-count agreement by interval / layer / tile)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import inputs, make_pair, noise_ensemble, rel_err
import odevio_b200
from oracle.pose_odernn import default_opt

dev = torch.device("cuda:0")
for B, bias in ((1200, 0.0), (256, 0.05)):
    ref, mod_tc = make_pair(dev, ode_solver="dopri5", ode_rtol=1e-3, ode_precision="tf32x3", bias_std=bias)
    mod_f = odevio_b200.PoseODERNN(default_opt(ode_solver="dopri5", ode_rtol=1e-3))
    mod_f.load_state_dict(ref.state_dict()); mod_f = mod_f.to(dev).eval()
    fv, fi, ts = inputs(B, irregular=True, seed=1)
    with torch.no_grad():
        p_tc, h_tc = mod_tc(fv.to(dev), fi.to(dev), ts.to(dev))
        p_f, h_f = mod_f(fv.to(dev), fi.to(dev), ts.to(dev))
    torch.cuda.synchronize()
    s_tc, s_f = mod_tc.last_stats.cpu().long(), mod_f.last_stats.cpu().long()
    same = (s_tc == s_f).all(-1)
    print(f"B={B} bias={bias}: same {same.float().mean():.4f} pose err tc-vs-fma {rel_err(p_tc.cpu(), p_f.cpu()):.2e}")
    print("  by interval:", [round(x, 3) for x in same.float().mean(dim=(1, 2)).tolist()])
    print("  by layer:", same.float().mean(dim=(0, 2)).tolist())
    print("  mean steps tc", s_tc[..., 0].float().mean().item(), "fma", s_f[..., 0].float().mean().item())
    print("  mean acc tc", s_tc[..., 1].float().mean().item(), "fma", s_f[..., 1].float().mean().item())
    rows = torch.randperm(B, generator=torch.Generator().manual_seed(3))[:24]
    with torch.no_grad():
        p_ref, h_ref = ref(fv[rows], fi[rows], ts[rows])
    stable, sp, sh = noise_ensemble(ref, fv[rows], fi[rows], ts[rows], n_members=6)
    ns, na = ref.last_stats["n_steps"], ref.last_stats["n_accepted"]
    for name, st, p in (("tc", s_tc, p_tc), ("fma", s_f, p_f)):
        neq = (st[:, :, rows, 0] != ns) | (st[:, :, rows, 1] != na)
        print(f"  {name} vs oracle(24 rows): mismatch {int(neq.sum())}/{neq.numel()} stable-mismatch {int((neq & stable).sum())} "
              f"unstable {int((~stable).sum())} pose err {rel_err(p.cpu()[rows], p_ref):.2e} spread {sp:.2e}")
    # timing
    for name, m in (("tc", mod_tc), ("fma", mod_f)):
        fvd, fid, tsd = fv.to(dev), fi.to(dev), ts.to(dev)
        with torch.no_grad():
            m(fvd, fid, tsd)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3): m(fvd, fid, tsd)
            torch.cuda.synchronize()
        print(f"  {name}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per forward")
