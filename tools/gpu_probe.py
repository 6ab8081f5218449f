"""Developer probe (not part of the product): run a few regressor configs on the GPU, print
parity against the oracle and CUDA-event timings.  Usage: python tools/gpu_probe.py [case ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_pair, inputs, run_pair

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)

def parity(name, B, S, irregular=True, **over):
    ref, mod = make_pair(dev, bias_std=0.05, **over)
    out = run_pair(ref, mod, *inputs(B, S, irregular=irregular), ensemble=3 if B <= 64 else 1)
    st = mod.last_stats.cpu()
    print(f"[parity] {name}: pose_err={out['pose_err']:.3e} (oracle spread {out['spread_pose']:.3e}) h_err={out['h_err']:.3e} ({out['spread_h']:.3e}) "
          f"steps_equal={out.get('steps_equal')} mismatch_entries={out.get('n_mismatch_entries')}/{out.get('n_entries')} "
          f"unstable={out.get('n_unstable_entries')} mismatch_stable={out.get('n_mismatch_stable_entries')} "
          f"status={out.get('status_max')} mean_steps={st[...,0].float().mean():.2f}", flush=True)

def timing(name, B, S=10, iters=3, **over):
    ref, mod = make_pair(dev, bias_std=0.05, **over)
    fv, fi, ts = inputs(B, S, irregular=True)
    fv, fi, ts = fv.to(dev), fi.to(dev), ts.to(dev)
    with torch.no_grad():
        mod(fv, fi, ts); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): mod(fv, fi, ts)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    st = mod.last_stats.cpu().float()
    steps = st[..., 0]
    # algorithmic flops
    D, H, n, L = 768, over.get("ode_hidden_dim", 512), over.get("ode_fn_num_layers", 3), over.get("rnn_num_layers", 2)
    F_ode = 2 * (D * H + (n - 1) * H * H + H * D)
    solver = over.get("ode_solver", "dopri5")
    if solver in ("rk4", "rk4_38"):
        evals = 4.0 * steps.sum().item()
    else:
        evals = (6.0 * steps + (steps > 0).float()).sum().item()
    g = 3 if over.get("ode_rnn_type", "rnn") == "gru" else 1
    flops = evals * F_ode + B * S * (L * 2 * 2 * D * D * g + 2 * (D * 128 + 128 * 6))
    print(f"[timing] {name}: B={B} {ms:.2f} ms/fwd  {B*S/ms*1e3:.0f} seq-steps/s  mean_steps={steps.mean():.2f} max={steps.max():.0f} "
          f"alg {flops/1e9:.1f} GFLOP -> {flops/ms/1e9:.2f} TFLOP/s", flush=True)

cases = sys.argv[1:] or ["tiny", "parity", "timing"]
if "tiny" in cases:
    parity("tiny rk4 B=8 S=2", 8, 2, irregular=False, ode_solver="rk4")
    parity("tiny dopri5 B=8 S=2", 8, 2)
if "parity" in cases:
    parity("rk4 B=16", 16, 10, irregular=False, ode_solver="rk4")
    parity("dopri5 B=16", 16, 10)
    parity("dopri5 rtol1e-3 B=64", 64, 10, ode_rtol=1e-3)
    parity("tsit5 B=16", 16, 10, ode_solver="tsit5")
    parity("heun B=16", 16, 10, ode_solver="heun")
    parity("gru B=16", 16, 10, ode_rnn_type="gru")
    parity("L3 H1024 n2 B=11", 11, 4, rnn_num_layers=3, ode_hidden_dim=1024, ode_fn_num_layers=2)
    parity("rows16 B=24", 24, 4, ode_rows_per_tile=16)
    parity("rows4 B=24", 24, 4, ode_rows_per_tile=4)
    parity("literal dopri5 B=16", 16, 10, ode_endpoint="dense", ode_exact_landing=False)
    parity("dopri5 rtol1e-3 B=256", 256, 10, ode_rtol=1e-3)
if "timing" in cases:
    timing("rk4 B=16", 16, ode_solver="rk4")
    timing("dopri5 rtol1e-3 B=1024 rt8", 1024, ode_rtol=1e-3, ode_rows_per_tile=8)
    timing("dopri5 rtol1e-3 B=1024 rt16", 1024, ode_rtol=1e-3, ode_rows_per_tile=16)
    timing("dopri5 rtol1e-3 B=4096 rt16", 4096, ode_rtol=1e-3, ode_rows_per_tile=16)
    timing("rk4 B=1024 rt8", 1024, ode_solver="rk4", ode_rows_per_tile=8)
