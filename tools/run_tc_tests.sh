set -x
cd /root/repo
timeout 300 python tools/tc_wide_ncu_case.py > gpurun_out/wide_plain.log 2>&1 || exit 1
tail -1 gpurun_out/wide_plain.log | cut -c1-250
timeout 900 ncu --set full --clock-control none --import-source on -k regex:odernn_tc_evolve -s 12 -c 1 -o gpurun_out/prof_tc_wide_r01 -f python tools/tc_wide_ncu_case.py > gpurun_out/ncu_wide.log 2>&1
ls -la gpurun_out/prof_tc_wide_r01.ncu-rep
