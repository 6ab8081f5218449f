set -x
cd /root/repo
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tc_final.json 2> gpurun_out/bench_tc_final.err; cut -c1-250 gpurun_out/bench_tc_final.json
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 > gpurun_out/bench_fp32_final.json 2> gpurun_out/bench_fp32_final.err; cut -c1-250 gpurun_out/bench_fp32_final.json
timeout 900 python bench.py --workload odernn_train --train-batch 4096 --steps 3 --warmup 1 > gpurun_out/bench_train_final.json 2> gpurun_out/bench_train_final.err; cut -c1-250 gpurun_out/bench_train_final.json
