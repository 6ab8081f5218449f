set -x
cd /root/repo
timeout 1200 python tools/sweep.py --out gpurun_out/sweep_fwd_tc.json > gpurun_out/sweep_tc.log 2>&1
tail -42 gpurun_out/sweep_tc.log | cut -c1-260
