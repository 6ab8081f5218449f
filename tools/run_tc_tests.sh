set -x
cd /root/repo
timeout 1200 python -m pytest tests/test_odernn_gpu.py tests/test_golden_gpu.py tests/test_odernn_tc_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --precision fp32 2> gpurun_out/bench_fp32.err | cut -c1-330
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_tc.err > gpurun_out/bench_tc.json; cut -c1-330 gpurun_out/bench_tc.json
SUB=4 timeout 300 python tools/gpu_tc_timing.py 2>&1 | grep "rows=\(1920\|2048\)"
