set -x
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-her "${@:3}" > gpurun_out/bench_r1_$2_${N}gpu.json 2> gpurun_out/bench_${N}gpu_$2.err; }
run 29511 reach_joints
run 29512 reach_ee --control ee
run 29513 pick_and_place_ee --task pick_and_place --control ee --envs 32768
