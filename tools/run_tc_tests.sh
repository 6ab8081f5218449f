set -x
cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_fwd_tc_n2.json 2> gpurun_out/b2tc.err
tail -3 gpurun_out/b2tc.err; cut -c1-400 gpurun_out/bench_fwd_tc_n2.json
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
