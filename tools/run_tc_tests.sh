set -x
cd /root/repo
timeout 600 python tools/gpu_tc_probe.py 2>&1 | tail -40
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32x3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
tail -3 gpurun_out/bench_tc.err
cat gpurun_out/bench_tc.json
