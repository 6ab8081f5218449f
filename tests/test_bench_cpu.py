"""Host-side pieces of bench.py that run without a GPU: the clock sampler degrades to an explicit 'no clock source' record (it
must never hang or raise on a box without NVML / nvidia-smi), and the reference arm's JSON line keeps the contract's keys."""
import importlib.util
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_without_gpu_is_explicit_and_quick():
    bench = _bench()
    t0 = time.time()
    s = bench.ClockSampler(0, None)
    s.start()
    s.begin()
    out = s.stop()
    assert time.time() - t0 < 10.0
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
    if out["samples"] == 0:
        assert out["sm_mhz"] is None


def test_clock_sampler_parses_helper_lines_inside_the_timed_region_only():
    bench = _bench()
    s = bench.ClockSampler(0, None)
    s.kind, s.proc = "nvml", subprocess.Popen([sys.executable, "-c", "import time; time.sleep(30)"])
    now = time.time()
    s.lines = [(now - 5.0, f"{now - 5.0} 1200 1965 "),                       # before begin(): ignored
               (now + 0.01, f"{now + 0.01} 1965 1965 "),
               (now + 0.02, f"{now + 0.02} 1950 1965 sw_power_cap"),
               (now + 0.03, "garbage line")]
    s.t0 = now
    time.sleep(0.05)
    out = s.stop()
    assert out["samples"] == 2 and out["sm_max_mhz"] == 1965.0 and out["sm_mhz"] == 1957.5
    assert out["reasons"] == ["sw_power_cap"]


def test_reference_arm_line_has_the_contract_keys(tmp_path):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    code = ("import sys, bench; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; "
            "bench.WORKLOAD['B'] = 8; bench.main()")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["higher_is_better"] is True
    assert d["metric"] == "integrated_sequence_steps_per_sec" and d["unit"] == "sequence-steps/s"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
