import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_pair, inputs, run_pair, rel_err
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200)

def dbg(name, B, S, irregular=True, bias_std=0.05, **over):
    ref, mod = make_pair(dev, bias_std=bias_std, **over)
    fv, fi, ts = inputs(B, S, irregular=irregular)
    out = run_pair(ref, mod, fv, fi, ts)
    st = mod.last_stats.cpu().long()
    rs, ra = ref.last_stats["n_steps"], ref.last_stats["n_accepted"]
    neq = (st[..., 0] != rs) | (st[..., 1] != ra)
    print(f"== {name}: pose_err={out['pose_err']:.3e} h_err={out['h_err']:.3e} mismatches={int(neq.sum())}/{neq.numel()}")
    perr = (out['pose'] - out['pose_ref']).abs().amax(-1) / out['pose_ref'].abs().max()
    print(" pose err per interval (max over rows):", ["%.1e" % v for v in perr.amax(0).tolist()])
    print(" mismatch per interval:", neq.sum((1, 2)).tolist(), " per layer:", neq.sum((0, 2)).tolist())
    idx = neq.nonzero()[:12]
    for i, l, b in idx.tolist():
        print(f"   (i={i},l={l},b={b}) gpu steps/acc={st[i,l,b,0].item()}/{st[i,l,b,1].item()} ref={rs[i,l,b].item()}/{ra[i,l,b].item()} gap={ts[b,i+1]-ts[b,i]:.4f} t0={ts[b,i]-ts[b,0]:.4f}")
    print(" gpu steps[i=0]:", st[0, :, :, 0].tolist())
    print(" ref steps[i=0]:", rs[0].tolist())

#dbg("dopri5 B=8 S=3 bias0 regular", 8, 3, irregular=False, bias_std=0.0)
#dbg("dopri5 B=8 S=3 bias regular", 8, 3, irregular=False)
#dbg("dopri5 B=8 S=3 bias irregular", 8, 3)
#dbg("dopri5 B=24 S=4 (8e-5 case)", 24, 4)
#dbg("dopri5 L=1 B=8 S=3", 8, 3, rnn_num_layers=1)
#dbg("L3 H1024 n2 B=11", 11, 4, rnn_num_layers=3, ode_hidden_dim=1024, ode_fn_num_layers=2)

def trace(name, B, S, irregular=True, bias_std=0.05, T=6, **over):
    ref, mod = make_pair(dev, bias_std=bias_std, ode_trace_steps=T, **over)
    fv, fi, ts = inputs(B, S, irregular=irregular)
    out = run_pair(ref, mod, fv, fi, ts)
    tg = mod.last_trace.cpu(); tr = ref.last_stats["trace"]
    st = mod.last_stats.cpu().long(); rs = ref.last_stats["n_steps"]
    print(f"== trace {name}: pose_err={out['pose_err']:.3e}")
    for (i, l, b) in [(0, 0, 0), (0, 1, 1), (1, 0, 0), (2, 1, 3)]:
        if i >= S or b >= B: continue
        print(f" (i={i},l={l},b={b}) steps gpu={st[i,l,b,0].item()} ref={rs[i,l,b].item()}")
        print("   gpu dt   :", ["%.6e" % v for v in tg[i, l, b, :, 0].tolist()])
        print("   ref dt   :", ["%.6e" % v for v in tr[i, l, b, :, 0].tolist()])
        print("   gpu ratio:", ["%.4e" % v for v in tg[i, l, b, :, 1].tolist()])
        print("   ref ratio:", ["%.4e" % v for v in tr[i, l, b, :, 1].tolist()])
    both = (tg[..., 1] > 0) & (tr[..., 1] > 0)
    rel = ((tg[..., 1] - tr[..., 1]).abs() / tr[..., 1].clamp_min(1e-30))[both]
    print(" ratio rel diff: median %.2e  p90 %.2e  max %.2e  (n=%d)" % (rel.median(), rel.quantile(0.9), rel.max(), rel.numel()))
    big = tr[..., 1] > 1e-3
    relb = ((tg[..., 1] - tr[..., 1]).abs() / tr[..., 1].clamp_min(1e-30))[both & big]
    if relb.numel(): print(" ratio rel diff where ref ratio>1e-3: median %.2e max %.2e (n=%d)" % (relb.median(), relb.max(), relb.numel()))

trace("dopri5 B=8 S=3 regular", 8, 3, irregular=False)
trace("L3 H1024 n2 B=11", 11, 4, rnn_num_layers=3, ode_hidden_dim=1024, ode_fn_num_layers=2)
