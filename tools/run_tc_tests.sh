set -x
cd /root/repo
timeout 900 python -m pytest tests/test_odernn_tc_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
tail -3 gpurun_out/bench_tc.err
cut -c1-330 gpurun_out/bench_tc.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tc.json')); r=d['roofline']
print({k:r[k] for k in ('achieved','launch_ms','kernel_share_of_step','rows_in_cluster_kernel','rows_in_ffma_side_launch')})
PY
