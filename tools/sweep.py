#!/usr/bin/env python
"""BASELINE configs[4]: inference throughput sweep of the fused PoseODERNN forward -- batch 256 -> 65536
sequences x ODEFunc hidden 128 -> 1024, fixed-step rk4 vs dopri5 (rtol 1e-3), one GPU per process
(rows are independent: N GPUs = N such processes on row shards, see bench.py --gpus).

    python tools/sweep.py [--out profiles/r01_sweep_fwd.json] [--quick]

Prints one JSON line per point: sequence-steps/s (CUDA events, 1 warm-up + 2 timed forwards, inputs
resident in HBM) and algorithmic TFLOP/s from the kernel's own step statistics."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import odevio_b200
from odevio_b200 import synth
from bench import init_like_deepvio
from types import SimpleNamespace


def point(dev, B, H, solver, S=10, L=2, n=3, reps=2, precision="tf32x3"):
    opt = SimpleNamespace(v_f_len=512, i_f_len=256, fuse_method="cat", ode_hidden_dim=H, ode_fn_num_layers=n,
                          ode_activation_fn="tanh", ode_solver=solver, ode_rnn_type="rnn", rnn_num_layers=L,
                          rnn_hidden_dim=1024, rnn_dropout_out=0.0, ode_rtol=1e-3, ode_atol=1e-6, ode_dt0=1e-4,
                          ode_precision=precision)
    model = odevio_b200.PoseODERNN(opt)
    init_like_deepvio(model, seed=0)
    model = model.to(dev).eval()
    fv, fi = synth.features(B, S, 512, 256, seed=0)
    ts = synth.timestamps(B, S, irregular=True, seed=0)
    fv, fi, ts = fv.to(dev), fi.to(dev), ts.to(dev)
    with torch.no_grad():
        model(fv, fi, ts)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            model(fv, fi, ts)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    st = model.last_stats[..., 0].double()
    D = 768
    f_ode = 2 * (D * H + (n - 1) * H * H + H * D)
    evals = (4.0 * st.sum().item()) if solver == "rk4" else (6.0 * st + (st > 0).double()).sum().item()
    flops = evals * f_ode + B * S * (L * 4 * D * D + 2 * (D * 128 + 128 * 6))
    assert int(model.last_status.max().item()) == 0
    return {"B": B, "H": H, "solver": solver, "precision": precision, "ms_per_forward": ms, "seq_steps_per_s": B * S / (ms * 1e-3),
            "algorithmic_tflops": flops / (ms * 1e-3) / 1e12, "mean_steps_per_interval": st.mean().item()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_fwd.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3"],
                    help="tf32x3: tensor-core solver (clusters of 8 up to 16 tiles, clusters of 4 beyond); fp32: FFMA kernel")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    Bs = [256, 4096] if args.quick else [256, 1024, 4096, 16384, 65536]
    Hs = [128, 512] if args.quick else [128, 256, 512, 1024]
    rows = []
    for solver in ("rk4", "dopri5"):
        for H in Hs:
            for B in Bs:
                try:
                    r = point(dev, B, H, solver, precision=args.precision)
                except odevio_b200.OdevioError as err:       # shape outside the tcgen05 tiling: FFMA kernel
                    r = point(dev, B, H, solver, precision="fp32")
                    r["note"] = f"{args.precision} unsupported here ({err}); FFMA kernel"
                rows.append(r)
                print(json.dumps(r), flush=True)
    with open(args.out, "w") as fh:
        json.dump({"gpu": torch.cuda.get_device_name(0), "points": rows}, fh, indent=1)


if __name__ == "__main__":
    main()
