# Scratch runner for `gpurun`: full validation of the current build on one B200 (edit freely).
set -x
cd /root/repo
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_tc_final.json 2> gpurun_out/bench_tc_final.err; cut -c1-250 gpurun_out/bench_tc_final.json
