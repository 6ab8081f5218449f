cd /root/repo
timeout 300 python tools/gpu_tc_timeline.py 1920 2>&1 | tail -12
