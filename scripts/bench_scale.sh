N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-her > gpurun_out/bench_r1_reach_joints_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
