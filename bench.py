#!/usr/bin/env python
"""Benchmark of the fused PoseODERNN forward (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (PoseODERNN.forward behind the C ABI) over one batch of
synthetic KITTI-shaped fused features: B = 1024 sequences per GPU x S = 10 observation intervals
= 10240 integrated sequence-steps per GPU per step, irregular timestamps (frame drop p ~ U[0,0.5],
src/data/KITTI_dataset.py:63-74), dopri5 rtol=1e-3 atol=1e-6 dt0=1e-4, D=768, H=512, n=3, L=2, nn.RNN.

Prints ONE JSON line (rank 0): metric = integrated sequence-steps/s, `value` with inputs resident
in HBM (CUDA events, max over ranks), `e2e` through the public module API from pinned HOST buffers
with H2D/D2H inside the timed region, `roofline` for the dominant (fused) kernel, `cpu_baseline` =
the oracle restatement of the reference's CPU regressor path on a bounded sample.

`--impl reference` times the reference arm: the oracle restatement of the reference's CPU path
(torchode/torchcde/torchdiffeq are not installable offline) on the host cores.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(B=1024, S=10, v_f_len=512, i_f_len=256, H=512, n=3, L=2, rnn="rnn",
                solver="dopri5", rtol=1e-3, atol=1e-6, dt0=1e-4)
CPU_SAMPLE_B = 1024         # the reference arm and cpu_baseline run the SAME B = 1024 workload on the host cores
METRIC = "integrated_sequence_steps_per_sec"
UNIT = "sequence-steps/s"


def workload_name(b=None):
    w = WORKLOAD
    return (f"PoseODERNN forward, irregular ts (frame drop 0-50%), {w['solver']} rtol={w['rtol']:g} "
            f"atol={w['atol']:g} dt0={w['dt0']:g}, B={b or w['B']}/GPU x S={w['S']}, D={w['v_f_len'] + w['i_f_len']}, "
            f"H={w['H']}, n={w['n']}, L={w['L']} nn.RNN, random-init (DeepVIO rule), BASELINE configs[1]")


# set by --precision: "fp16x3" (default: odernn_h3.cu, ONE tcgen05 cluster kernel per forward, 3xFP16 = fp32-accurate) |
# "tf32x3" (round-1 tensor-core solver, per-interval launches) | "fp32" (CUDA-core FFMA kernel)
PRECISION = "fp16x3"


def make_opt():
    from types import SimpleNamespace
    w = WORKLOAD
    return SimpleNamespace(v_f_len=w["v_f_len"], i_f_len=w["i_f_len"], fuse_method="cat",
                           ode_hidden_dim=w["H"], ode_fn_num_layers=w["n"], ode_activation_fn="tanh",
                           ode_solver=w["solver"], ode_rnn_type=w["rnn"], rnn_num_layers=w["L"],
                           rnn_hidden_dim=1024, rnn_dropout_out=0.0, ode_rtol=w["rtol"], ode_atol=w["atol"],
                           ode_dt0=w["dt0"], ode_precision=PRECISION)


def init_like_deepvio(model, seed=0):
    """Reference init rule (src/models/DeepVIO.py:77-87): kaiming-normal Linear, zero bias; RNN default."""
    import torch
    import torch.nn as nn
    torch.manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.Linear):
            nn.init.kaiming_normal_(m.weight.data)
            m.bias.data.zero_()
    for name, p in model.named_parameters():       # deterministic RNN weights for every rank
        if name.startswith("rnn."):
            nn.init.uniform_(p.data, -model.f_len ** -0.5, model.f_len ** -0.5)


def algorithmic_flops(stats, B, S):
    """sum over (interval, layer, row) of (1 + 6*n_steps) * F_ode + jump + head."""
    w = WORKLOAD
    D, H, n, L = w["v_f_len"] + w["i_f_len"], w["H"], w["n"], w["L"]
    f_ode = 2 * (D * H + (n - 1) * H * H + H * D)
    f_rnn = L * 2 * (2 * D * D)
    f_reg = 2 * (D * 128 + 128 * 6)
    steps = stats[..., 0].double()
    evals = (6.0 * steps + (steps > 0).double()).sum().item()
    return evals * f_ode + B * S * (f_rnn + f_reg), evals


def _gpu_uuid(dev):
    try:
        import torch
        return "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        return None


_NVML_HELPER = r"""
import sys, time
import pynvml as n
n.nvmlInit()
uuid, index = sys.argv[1], int(sys.argv[2])
h = None
if uuid:
    for u in (uuid, uuid.encode()):
        try:
            h = n.nvmlDeviceGetHandleByUUID(u); break
        except Exception:
            h = None
if h is None:
    h = n.nvmlDeviceGetHandleByIndex(index)
mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
bits = [("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
        ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)]
while True:
    sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
    try:
        r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    print(time.time(), sm, mx, "|".join(nm for nm, b in bits if r & b), flush=True)
    time.sleep(0.005)
"""


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region.  A helper PROCESS polls NVML every 5 ms (started well
    before the region, so even a 0.2 s region gets tens of samples; a separate process, so that a profiler attached to this
    one cannot stall the bench through it); `nvidia-smi -lms 100` when the helper cannot load NVML.  start() launches the
    helper, begin() / stop() bracket the timed region: only samples stamped inside it count."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.lines, self.proc, self.kind, self.t0 = index, uuid, [], None, None, None

    def _spawn(self, cmd):
        self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def start(self):
        try:
            self._spawn([sys.executable, "-c", _NVML_HELPER, self.uuid or "", str(self.index)])
            self.kind = "nvml"
            deadline = time.time() + 3.0                 # the helper either prints within moments or died (no pynvml / NVML)
            while time.time() < deadline and not self.lines and self.proc.poll() is None:
                time.sleep(0.01)
            if self.lines:
                return
            self.proc.kill()
        except OSError:
            pass
        try:
            self.lines = []
            self._spawn(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"])
            self.kind = "smi"
        except OSError:
            self.proc, self.kind = None, None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def begin(self):
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (pynvml and nvidia-smi unavailable)"], "samples": 0}
        t1 = time.time()
        if self.kind == "smi":
            time.sleep(0.15)
        self.proc.terminate()
        t0 = self.t0 if self.t0 is not None else 0.0
        sm, mx, reasons = [], [], set()
        for stamp, ln in list(self.lines):
            try:
                if self.kind == "nvml":
                    f = ln.split(" ")
                    ts_, s_, m_ = float(f[0]), float(f[1]), float(f[2])
                    rs = [x for x in (f[3].split("|") if len(f) > 3 else []) if x]
                else:
                    f = [x.strip() for x in ln.split(",")]
                    ts_, s_, m_ = stamp, float(f[1]), float(f[2])
                    rs = [nm for nm, val in zip(self.NAMES, f[5:9]) if val.lower().startswith("active")]
            except (ValueError, IndexError):
                continue
            if ts_ < t0 or ts_ > t1 + 0.2:
                continue
            sm.append(s_); mx.append(m_); reasons.update(rs)
        load = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml helper process, 5 ms poll" if self.kind == "nvml" else "nvidia-smi -lms 100"}


def cpu_oracle_rate(steps, warmup, B, opt=None, irregular=True):
    """Oracle restatement of the reference CPU path on `B` sequences of the workload."""
    import torch
    from oracle.modules import deepvio_initialization
    from oracle.pose_odernn import OraclePoseODERNN
    from odevio_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    ref = OraclePoseODERNN(opt or make_opt())
    deepvio_initialization(ref)
    ref.eval()
    w = WORKLOAD
    fv, fi = synth.features(B, w["S"], w["v_f_len"], w["i_f_len"], seed=0)
    ts = synth.timestamps(B, w["S"], irregular=irregular, seed=0)
    times = []
    with torch.no_grad():
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            ref(fv, fi, ts)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return B * w["S"] / sec, sec, torch.get_num_threads()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, rank):
    if rank != 0:
        return
    rate, sec, threads = cpu_oracle_rate(args.steps, args.warmup, CPU_SAMPLE_B)
    sample = (f"all {CPU_SAMPLE_B} sequences of the workload per step, oracle restatement of the reference "
              f"CPU regressor path (torchode unavailable offline), eager PyTorch fp32, {cpu_model()}")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(CPU_SAMPLE_B), "device": "host CPU"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def ffma_peak_tflops(torch, lib, dev):
    """This GPU's fp32 FMA peak from the library's dense-FFMA microbenchmark (CUDA events)."""
    import ctypes as C
    sink = torch.zeros(4, device=dev)
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    flops = C.c_double(0)
    stream = torch.cuda.current_stream(dev).cuda_stream
    best = 0.0
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.odevio_microbench_ffma(20000, nsm * 2 * 4, C.c_void_p(sink.data_ptr()), C.byref(flops),
                                        C.c_void_p(stream))
        e1.record()
        torch.cuda.synchronize(dev)
        if rc != 0:
            return None
        if it:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import odevio_b200
    from odevio_b200 import _lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the odevio_b200 path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOAD
    B, S = w["B"], w["S"]
    model = odevio_b200.PoseODERNN(make_opt())
    init_like_deepvio(model, seed=0)
    model = model.to(dev).eval()
    # each rank integrates its own shard of independent sequences (no data-path collective)
    fv_h, fi_h = synth.features(B, S, w["v_f_len"], w["i_f_len"], seed=rank)
    ts_h = synth.timestamps(B, S, irregular=True, seed=rank)
    fv_h, fi_h, ts_h = fv_h.pin_memory(), fi_h.pin_memory(), ts_h.pin_memory()
    fv, fi, ts = fv_h.to(dev), fi_h.to(dev), ts_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing: K steps, L2 flushed between timed iterations
    with torch.no_grad():
        sampler = ClockSampler(local_rank, _gpu_uuid(dev))
        if rank == 0:
            sampler.start()                      # helper process up and polling before the warm-up ends
        for _ in range(max(args.warmup, 3)):
            model(fv, fi, ts)
        barrier()
        sampler.begin()
        evs = []
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model(fv, fi, ts)
            e1.record()
            evs.append((e0, e1))
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        step_ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
        stats = model.last_stats.clone()
        status = int(model.last_status.max().item())

        # ---- dominant kernel's average launch duration, live: CUDA events around every solver launch on its stream
        tc_kernel_ms, tc_launches, tc_geo = None, 0, None
        if PRECISION in ("tf32x3", "fp16x3"):
            import ctypes as C
            lib.odevio_debug_tc_timing(1, None, None)
            for _ in range(args.steps):
                flush.fill_(1)
                model(fv, fi, ts)
            torch.cuda.synchronize(dev)
            tot, cnt = C.c_float(0), C.c_int32(0)
            lib.odevio_debug_tc_timing(-1, C.byref(tot), C.byref(cnt))
            lib.odevio_debug_tc_timing(0, None, None)
            geo = (C.c_int32 * 3)()
            lib.odevio_debug_tc_geometry(geo)
            tc_kernel_ms, tc_launches, tc_geo = tot.value, cnt.value, list(geo)

        # ---- end to end through the public API from pinned host buffers
        pose_h = torch.empty(B, S, 6, dtype=torch.float32).pin_memory()
        for _ in range(2):
            p, _h = model(fv_h.to(dev, non_blocking=True), fi_h.to(dev, non_blocking=True),
                          ts_h.to(dev, non_blocking=True))
            pose_h.copy_(p, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            p, _h = model(fv_h.to(dev, non_blocking=True), fi_h.to(dev, non_blocking=True),
                          ts_h.to(dev, non_blocking=True))
            pose_h.copy_(p, non_blocking=True)
            torch.cuda.synchronize(dev)              # the caller consumes the poses every step
        barrier()
        e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = total_ms.item(), e2e_ms.item()
    train_info = None
    if args.train:
        del flush
        torch.cuda.empty_cache()
        train_info = measure_train(args, rank, world, dev)
    if status != 0:
        raise SystemExit(f"bench.py: solver status {status} (non-finite norm / max_steps) -- result invalid")

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * B * S / (ms_per_step * 1e-3)
        e2e_value = world * B * S / (e2e_ms / args.steps * 1e-3)
        flops, evals = algorithmic_flops(stats.cpu(), B, S)
        achieved_tf = flops / (ms_per_step * 1e-3) / 1e12
        peaks, peak_src = measured_peaks()
        fma_peak = ffma_peak_tflops(torch, lib, dev)
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            with open(prof) as fh:
                traffic = json.load(fh).get({"tf32x3": "dram_bytes_per_launch_tc", "fp16x3": "dram_bytes_per_launch_h3"}.get(
                    PRECISION, "dram_bytes_per_launch"))
        n_side = 0
        if PRECISION == "fp16x3":
            # ONE launch per forward: solver loops of all S intervals + rnn jump + pose head in odernn_h3_kernel
            per_launch_ms = tc_kernel_ms / tc_launches
            launches_per_step = tc_launches / args.steps
            h3_achieved = flops / launches_per_step / (per_launch_ms * 1e-3) / 1e12
            fp16x3_peak = peaks["bf16_tflops_sustained"] / 3.0
            roofline = {
                "bound": "tensor", "kernel": "odernn_h3_kernel<64>",
                "achieved": h3_achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": h3_achieved / peaks["bf16_tflops_sustained"], "peak_source": peak_src + ", sustained bf16",
                "traffic": traffic,
                "algorithmic_flops_per_launch": flops / launches_per_step, "launch_ms": per_launch_ms,
                "launches_per_step": launches_per_step, "kernel_share_of_step": tc_kernel_ms / args.steps / ms_per_step,
                "clusters": tc_geo[0], "max_coresident_clusters": tc_geo[1], "rows": tc_geo[2],
                "vector_field_evals_per_launch": evals / launches_per_step,
                "tensor_3xfp16": {"achieved": h3_achieved, "peak": fp16x3_peak, "frac": h3_achieved / fp16x3_peak,
                                  "unit": "TFLOP/s",
                                  "note": "fp32 parity needs 3 fp16 MMAs per product (hi*hi, lo*hi, hi*lo): the fp32-accurate "
                                          "tensor ceiling is the fp16/bf16 peak / 3"},
                "note": "launch duration measured live with CUDA events around the solver launch on its stream "
                        "(odevio_debug_tc_timing) in extra un-profiled forwards after the timed region; algorithmic FLOPs "
                        "from the kernel's own step statistics (SURVEY.md 8d)",
            }
        elif PRECISION == "tf32x3":
            # the cluster kernel integrates rows g = l * B + b < tc_rows of every interval; the rest (side launch) is FFMA
            D_, H_, n_ = w["v_f_len"] + w["i_f_len"], w["H"], w["n"]
            f_ode = 2 * (D_ * H_ + (n_ - 1) * H_ * H_ + H_ * D_)
            tc_rows = tc_geo[2]
            n_side_seq = B - tc_rows // w["L"]
            n_side = 1 if n_side_seq > 0 else 0
            st_all = stats.cpu()[..., 0].double()                       # [S, L, B]
            in_tc = torch.ones(S, B, dtype=torch.bool)
            if n_side_seq > 0:
                # the library's per-interval rule (tc_select_kernel): the n_side sequences with the shortest interval
                # (ties by index) run in the FFMA side launch
                gaps = (ts_h[:, 1:] - ts_h[:, :-1]).float()
                for i in range(S):
                    in_tc[i, torch.argsort(gaps[:, i], stable=True)[:n_side_seq]] = False
            st_rows = st_all * in_tc[:, None, :]
            tc_evals = (6.0 * st_rows + (st_rows > 0).double()).sum().item()
            per_launch_ms = tc_kernel_ms / tc_launches
            tc_flops_per_launch = tc_evals * f_ode / S
            tc_achieved = tc_flops_per_launch / (per_launch_ms * 1e-3) / 1e12
            tf32x3_peak = peaks["bf16_tflops_sustained"] / 2.0 / 3.0
            roofline = {
                "bound": "tensor", "kernel": "odernn_tc_evolve_kernel<8,8>",
                "achieved": tc_achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": tc_achieved / peaks["bf16_tflops_sustained"], "peak_source": peak_src + ", sustained bf16",
                "traffic": traffic,
                "algorithmic_flops_per_launch": tc_flops_per_launch, "launch_ms": per_launch_ms,
                "launches_per_step": S, "kernel_share_of_step": tc_kernel_ms / args.steps / ms_per_step,
                "clusters": tc_geo[0], "max_coresident_clusters": tc_geo[1], "rows_in_cluster_kernel": tc_rows,
                "rows_in_ffma_side_launch": w["L"] * B - tc_rows,
                "whole_step": {"achieved": achieved_tf, "algorithmic_flops": flops, "vector_field_evals": evals},
                "tensor_3xtf32": {"achieved": tc_achieved, "peak": tf32x3_peak, "frac": tc_achieved / tf32x3_peak,
                                  "unit": "TFLOP/s",
                                  "note": "fp32 parity needs 3 TF32 MMAs per product (hi*hi, lo*hi, hi*lo) and TF32 runs at half "
                                          "the bf16 rate: the fp32-accurate tensor ceiling is bf16 peak / 6"},
                "note": "launch duration measured live with CUDA events around every solver launch on its stream "
                        "(odevio_debug_tc_timing) in extra un-profiled forwards after the timed region; latency-bound: "
                        "15 clusters x 8 CTAs each walk one 128-row tile through ~36 dependent ODEFunc evaluations per launch",
            }
        else:
            roofline = {
                "bound": "tensor", "kernel": "odernn_fwd_kernel<8,2>",
                "achieved": achieved_tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["bf16_tflops_sustained"], "peak_source": peak_src + ", sustained bf16",
                "traffic": traffic,
                "algorithmic_flops_per_launch": flops, "vector_field_evals_per_launch": evals,
                "note": "fp32 parity mode runs the GEMMs on CUDA-core FFMA, so the pipe that bounds it is fp32 FMA "
                        "(fma_fp32 below, peak measured live by odevio_microbench_ffma); the tensor-pipe fraction is "
                        "reported against the measured bf16 peak as the contract asks",
                "fma_fp32": {"achieved": achieved_tf, "peak": fma_peak,
                             "frac": (achieved_tf / fma_peak) if fma_peak else None, "unit": "TFLOP/s"},
            }
        n_prepack = (w["n"] + 1) + w["L"] * 3 + 1
        cpu_rate, cpu_sec, cpu_threads = cpu_oracle_rate(steps=4, warmup=1, B=CPU_SAMPLE_B) if world == 1 else (None, None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "global_batch": world * B,
                       "parallelism": f"independent sequences sharded over {world} GPU(s), no data-path collective",
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                       "mean_solver_steps_per_interval": stats[..., 0].float().mean().item(),
                       "precision": PRECISION + {"tf32x3": " (ODEFunc GEMMs on tcgen05 as 3xTF32, fp32-accurate; jump/head and the "
                                                           "rows beyond the co-resident clusters on FFMA)",
                                                 "fp16x3": " (every GEMM of the forward on tcgen05 as 3xFP16, fp32-accurate; one "
                                                           "cluster kernel per forward)"}.get(PRECISION, " (CUDA-core FFMA)")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": fv_h.numel() * 4 + fi_h.numel() * 4 + ts_h.numel() * 4,
                    "d2h_bytes_per_step": pose_h.numel() * 4, "ms_per_step": e2e_ms / args.steps},
            # kernels of this library launched inside the timed region (fp16x3: the weights are packed once, by the first
            # warm-up forward -- cfg.weights_prepacked -- so a timed step is exactly one launch)
            "gpu_launches": (args.steps if PRECISION == "fp16x3" else (n_prepack + 1) * args.steps if PRECISION == "fp32" else
                             (n_prepack + (w["n"] + 1) + S * (2 + n_side)) * args.steps),
            "roofline": roofline,
        }
        if cpu_rate is not None:
            line["cpu_baseline"] = {
                "value": cpu_rate, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                "sample": f"all {CPU_SAMPLE_B} sequences of the workload, 4 forwards after 1 warm-up ({cpu_sec:.2f} s each), "
                          f"oracle restatement of the reference CPU regressor path, eager PyTorch fp32, {cpu_model()}"}
            line["config0"] = config0_row(torch, odevio_b200, dev)
        if args.train:
            line["train"] = train_info
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def config0_row(torch, odevio_b200, dev):
    """BASELINE configs[0]: PoseODERNN forward, fixed-step rk4, B = 16 x seq_len 11, regular timestamps -- the reference's
    own CPU-runnable case: the CPU oracle port and this library (same --precision) side by side."""
    import copy
    from odevio_b200 import synth
    w = WORKLOAD
    opt = make_opt()
    opt.ode_solver = "rk4"
    cpu_rate, cpu_sec, threads = cpu_oracle_rate(steps=10, warmup=2, B=16, opt=copy.copy(opt), irregular=False)
    model = odevio_b200.PoseODERNN(opt)
    init_like_deepvio(model, seed=0)
    model = model.to(dev).eval()
    fv, fi = synth.features(16, w["S"], w["v_f_len"], w["i_f_len"], seed=0)
    ts = synth.timestamps(16, w["S"], irregular=False, seed=0)
    fv, fi, ts = fv.to(dev), fi.to(dev), ts.to(dev)
    with torch.no_grad():
        for _ in range(3):
            model(fv, fi, ts)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model(fv, fi, ts)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 10
    return {"workload": "PoseODERNN forward, fixed-step rk4, B=16 x seq_len 11 (S=10), regular ts, BASELINE configs[0]",
            "gpu_ms_per_step": ms, "gpu_value": 16 * w["S"] / (ms * 1e-3),
            "cpu_ms_per_step": cpu_sec * 1e3, "cpu_value": cpu_rate, "cpu_cores": threads, "unit": UNIT,
            "note": "one 64-row tile = 4 of 148 SMs busy: a latency measurement, not a throughput one"}


def measure_train(args, rank, world, dev):
    """BASELINE configs[3] next to the headline: PoseODERNN training step (forward with checkpoints + fused backward + ONE
    NCCL all-reduce of the flat Pose_net gradient bucket + clip + Adam as its epilogue: training.fused_train_step),
    global batch `--train-batch` strong-scaled over the ranks.  Collective on every rank; returns the dict rank 0 prints."""
    import torch
    import torch.distributed as dist
    import odevio_b200
    from odevio_b200 import distributed as D, synth, training
    w = WORKLOAD
    GB, S = args.train_batch, w["S"]
    info = {"workload": f"PoseODERNN training step (fwd + fused bwd + NCCL grad all-reduce + clip + Adam), global batch {GB} "
                        f"strong-scaled over {world} GPU(s), dopri5 rtol={w['rtol']:g}, BASELINE configs[3]",
            "global_batch": GB, "scaling": "strong", "n_gpus": world}
    try:
        a, b = D.shard_rows(GB, rank, world)
        opt_ns = make_opt()
        model = odevio_b200.PoseODERNN(opt_ns)
        init_like_deepvio(model, seed=0)
        model = model.to(dev).train()
        opt = training.FusedPoseNetAdam(model, lr=1e-4)
        fv, fi = synth.features(GB, S, w["v_f_len"], w["i_f_len"], seed=0)
        ts = synth.timestamps(GB, S, irregular=True, seed=0)
        gts = 0.01 * torch.randn(GB, S, 6, generator=torch.Generator().manual_seed(1))
        fv, fi, ts, gts = (t[a:b].to(dev) for t in (fv, fi, ts, gts))
        torch.cuda.reset_peak_memory_stats(dev)
        training.fused_train_step(model, opt, fv, fi, ts, gts, world_size=world)       # warm-up (allocations)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        steps = args.train_steps
        evs = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = training.fused_train_step(model, opt, fv, fi, ts, gts, world_size=world, events=evs)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        phases = {}
        for name, x0, x1 in evs:
            phases[name] = phases.get(name, 0.0) + x0.elapsed_time(x1) / steps
        vals = torch.tensor([e0.elapsed_time(e1) / steps, phases["allreduce"], phases["forward"], phases["loss_backward"],
                             phases["clip_adam"], torch.cuda.max_memory_allocated(dev) / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        per, ar, fw, bw, ad, mem = vals.tolist()
        info.update({"ms_per_step": per, "value": GB * S / (per * 1e-3), "unit": UNIT, "steps": steps, "warmup": 1,
                     "allreduce_ms": ar, "forward_ms": fw, "loss_backward_ms": bw, "clip_adam_ms": ad,
                     "allreduce_bytes": opt.flat.numel() * 4, "peak_mem_gb": mem, "loss": float(loss),
                     "timing": "CUDA events, max over ranks"})
        # the collective itself, both ways, on the same gradient bucket: NCCL all-reduce + the three-launch clip / Adam step
        # against ONE kernel over NVLink peer memory (odevio_allreduce_adam_peer: reduce-scatter by pull, clip, Adam on the
        # rank's slice, all-gather by push); its own try: a symmetric-memory problem must not cost the numbers above
        if world > 1:
            try:
                def timed(fn, n=20):
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize(dev); dist.barrier()
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record()
                    for _ in range(n):
                        fn()
                    t1.record(); torch.cuda.synchronize(dev)
                    return t0.elapsed_time(t1) / n * 1e3

                def nccl_step():
                    dist.all_reduce(opt.grads); opt.step()
                us_nccl = timed(nccl_step)
                peer = training.PeerFusedPoseNetAdam(model, lr=1e-4)
                peer.grads.copy_(opt.grads)
                us_peer = timed(peer.step_allreduce)
                both = torch.tensor([us_nccl, us_peer], dtype=torch.float64, device=dev)
                dist.all_reduce(both, op=dist.ReduceOp.MAX)
                info["collective"] = {"nccl_allreduce_plus_clip_adam_us": both[0].item(), "peer_one_kernel_us": both[1].item(),
                                      "bucket_bytes": opt.flat.numel() * 4, "launches": {"nccl": 4, "peer": 1},
                                      "note": "gradient all-reduce + clip + Adam of the 15 MB Pose_net bucket; max over ranks"}
                del peer
            except Exception as exc:
                info["collective"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        del model, opt, fv, fi, ts, gts
        torch.cuda.empty_cache()
    except Exception as exc:                       # the headline line must still be printed
        info["error"] = f"{type(exc).__name__}: {exc}"[:300]
    return info


def _device_setup(local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the odevio_b200 path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return dev


def run_train(args, rank, world, local_rank):
    """BASELINE configs[3]: PoseODERNN training step (forward + fused backward + NCCL gradient
    all-reduce + clip + Adam), global batch 4096 sharded over the ranks (strong scaling)."""
    import torch
    import torch.distributed as dist
    import odevio_b200
    from odevio_b200 import distributed as D, synth
    dev = _device_setup(local_rank, world)
    w = WORKLOAD
    GB, S = args.train_batch, w["S"]
    a, b = D.shard_rows(GB, rank, world)
    model = odevio_b200.PoseODERNN(make_opt())
    init_like_deepvio(model, seed=0)
    model = model.to(dev).train()
    opt = D.make_optimizer(model, lr=1e-4)
    fv, fi = synth.features(GB, S, w["v_f_len"], w["i_f_len"], seed=0)
    ts = synth.timestamps(GB, S, irregular=True, seed=0)
    gts = 0.01 * torch.randn(GB, S, 6, generator=torch.Generator().manual_seed(1))
    fv, fi, ts, gts = (t[a:b].to(dev) for t in (fv, fi, ts, gts))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 1)):
        D.train_step(model, opt, fv, fi, ts, gts, world_size=world)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = D.train_step(model, opt, fv, fi, ts, gts, world_size=world)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        per = ms.item() / args.steps
        print(json.dumps({
            "metric": "training_" + METRIC, "value": GB * S / (per * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": per, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "PoseODERNN training step (fwd + fused bwd + grad all-reduce + clip + Adam), "
                                   f"global batch {GB} sharded over {world} GPU(s), dopri5 rtol=1e-3, BASELINE configs[3]",
                       "global_batch": GB, "loss": float(loss)},
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cde(args, rank, world, local_rank):
    """BASELINE configs[2]: PoseCDE forward, B=1024, Hc = F = 128, cubic control path (and the
    reference's rectilinear-linear path), dopri5 atol=1e-6 rtol=1e-4.  Replicas only for N > 1
    (the batch-joint controller is not shard-invariant)."""
    import torch
    import odevio_b200
    from types import SimpleNamespace
    from odevio_b200 import synth
    dev = _device_setup(local_rank, 1)
    B, S, Hc = 1024, 10, 128
    out = {}
    for interp in ("cubic", "linear"):
        opt = SimpleNamespace(v_f_len=Hc // 2, i_f_len=Hc // 2, fuse_method="cat", cde_hidden_dim=Hc,
                              cde_fn_num_layers=3, cde_num_layers=3, cde_activation_fn="tanh", cde_solver="dopri5",
                              adjoint=False, cde_interp=interp)
        model = odevio_b200.PoseCDE(opt)
        torch.manual_seed(0)
        for m in model.modules():
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.kaiming_normal_(m.weight.data); m.bias.data.zero_()
        model = model.to(dev).train()
        fv, fi = synth.features(B, S, Hc // 2, Hc // 2, seed=rank)
        ts = synth.timestamps(B, S, irregular=True, seed=rank)
        fv, fi, ts = (0.2 * fv).to(dev), (0.2 * fi).to(dev), ts.to(dev)
        with torch.no_grad():
            for _ in range(max(args.warmup, 3)):
                model(fv, fi, ts)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                model(fv, fi, ts)
            e1.record()
            torch.cuda.synchronize(dev)
        per = e0.elapsed_time(e1) / args.steps
        st = model.last_stats.tolist()
        C = Hc + 1
        flops = st[2] * B * 2.0 * (3 * Hc * Hc + Hc * Hc * C + Hc * C)          # SURVEY.md 8d, per vf eval
        out[interp] = {"precision": model.last_precision, "ms_per_step": per, "seq_steps_per_s": B * S / (per * 1e-3), "solver_steps": st[0],
                       "accepted": st[1], "vf_evals": st[2], "status": st[3],
                       # nominal = every channel of the final Linear counted; the rectilinear path's
                       # time-only segments legitimately skip all but one channel group
                       ("algorithmic_tflops" if interp == "cubic" else "nominal_tflops_incl_skipped_channels"):
                           flops / (per * 1e-3) / 1e12}
        if args.train:
            # training step of the same workload: checkpointing forward + fused backward (odevio_cde_backward), the
            # reference's loss (scripts/train_model.py:72-77); gradients w.r.t. every parameter and the fused features
            g = torch.Generator().manual_seed(5)
            gts = (0.1 * torch.randn(B, S, 6, generator=g)).to(dev)
            fvg, fig = fv.clone().requires_grad_(True), fi.clone().requires_grad_(True)
            torch.cuda.reset_peak_memory_stats(dev)
            fw = bw = 0.0
            for it in range(1 + args.train_steps):
                model.zero_grad(set_to_none=True)
                fvg.grad = fig.grad = None
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
                pose, _ = model(fvg, fig, ts)
                ev[1].record()
                loss = 100 * torch.nn.functional.mse_loss(pose[:, :, :3], gts[:, :, :3]) + \
                    torch.nn.functional.mse_loss(pose[:, :, 3:], gts[:, :, 3:])
                loss.backward()
                ev[2].record()
                torch.cuda.synchronize(dev)
                if it > 0:
                    fw += ev[0].elapsed_time(ev[1]); bw += ev[1].elapsed_time(ev[2])
            n = max(args.train_steps, 1)
            out[interp]["train"] = {"forward_ckpt_ms": fw / n, "loss_backward_ms": bw / n, "ms_per_step": (fw + bw) / n,
                                    "seq_steps_per_s": B * S / ((fw + bw) / n * 1e-3), "steps": n, "warmup": 1,
                                    "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "loss": float(loss.item()),
                                    "backward_tflops": 3.0 * flops / (bw / n * 1e-3) / 1e12}
            del fvg, fig, pose, loss
            model._bwd_workspace = None
    if rank == 0:
        # roofline of the one cooperative launch of the cubic run (all of the forward is cde_fwd_kernel)
        from odevio_b200 import _lib
        peaks, peak_src = measured_peaks()
        fma_peak = ffma_peak_tflops(torch, _lib.load(), dev)
        ach = out["cubic"]["algorithmic_tflops"]
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            with open(prof) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch_cde")
        tc = out["cubic"]["precision"] == "fp16x3"
        roofline = {"bound": "tensor", "kernel": "cde_tc_kernel" if tc else "cde_fwd_kernel", "achieved": ach, "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"], "peak_source": peak_src + ", sustained bf16",
                    "traffic": traffic, "launch_ms": out["cubic"]["ms_per_step"], "launches_per_step": 1,
                    "algorithmic_flops_per_launch": ach * 1e12 * out["cubic"]["ms_per_step"] * 1e-3,
                    "fma_fp32": {"achieved": ach, "peak": fma_peak, "frac": (ach / fma_peak) if fma_peak else None, "unit": "TFLOP/s"},
                    "note": ("cde_tc_kernel: the final Hc -> Hc*(Hc+1) Linear runs on tcgen05 as 3xFP16 (3 MMAs per product: the "
                             "ceiling of the scheme is a third of the fp16 peak) with the weights resident in shared memory; "
                             "the Hc x Hc Linears of the row phase are CUDA-core FFMA" if tc else
                             "cde_fwd_kernel: every GEMM on CUDA-core FFMA; the fp32-FMA fraction is the pipe that bounds it")}
        print(json.dumps({"metric": "cde_" + METRIC, "value": out["cubic"]["seq_steps_per_s"], "unit": UNIT, "roofline": roofline,
                          "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": out["cubic"]["ms_per_step"], "higher_is_better": True, "scaling": "replicas only",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "PoseCDE forward, B=1024, S=10, Hc=F=128, n=3, dopri5 atol=1e-6 "
                                                 "rtol=1e-4, cubic (north_star) control path; 'linear' = reference "
                                                 "rectilinear path; BASELINE configs[2]"},
                          "detail": out}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="odernn_fwd", choices=["odernn_fwd", "odernn_train", "cde"],
                    help="odernn_fwd = the headline (BASELINE configs[1]); the others are extra measurements")
    ap.add_argument("--train-batch", type=int, default=4096)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--train-steps", type=int, default=2)
    ap.add_argument("--no-train", dest="train", action="store_false",
                    help="odernn_fwd: skip the extra `train` object (configs[3] training step incl. the NCCL gradient all-reduce)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=PRECISION, choices=["fp32", "tf32x3", "fp16x3"],
                    help="odernn_fwd: CUDA-core FFMA kernel, or a tcgen05 solver kernel (3xTF32 / 3xFP16, both fp32-accurate)")
    args = ap.parse_args()
    globals()["PRECISION"] = args.precision
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: for --gpus N > 1 launch with torch.distributed.run (one rank per GPU)")
    if args.workload == "odernn_train":
        run_train(args, rank, world, local_rank)
    elif args.workload == "cde":
        run_cde(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
