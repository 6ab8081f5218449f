cd /root/repo
timeout 1200 python tools/sweep.py --out gpurun_out/sweep_fwd_tc.json > gpurun_out/sweep_tc.log 2>&1
tail -3 gpurun_out/sweep_tc.log | cut -c1-220
timeout 600 python bench.py --workload cde --steps 5 --warmup 3 2>/dev/null | cut -c1-300
