set -x
cd /root/repo
timeout 600 python -m pytest tests/test_imu_encoder.py -x -q 2>&1 | tail -8
