set -x
cd /root/repo
timeout 300 python -m pytest tests/test_odefunc_gpu.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_odernn_tc_gpu.py -x -q 2>&1 | tail -3
SUB=4 timeout 300 python tools/gpu_tc_timing.py 2>&1 | grep "tf32x3 rows=\(128\|1920\|2048\)"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tc_op8.json 2> gpurun_out/bench_tc_op8.err; cut -c1-250 gpurun_out/bench_tc_op8.json
