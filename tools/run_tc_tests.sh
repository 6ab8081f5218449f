set -x
cd /root/repo
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tc_final.json 2> gpurun_out/bench_tc_final.err || exit 1
cut -c1-250 gpurun_out/bench_tc_final.json
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/plain_tc.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_tc1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:odernn_tc_evolve -s 12 -c 2 -o gpurun_out/prof_tc_r01 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_tc2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -1
