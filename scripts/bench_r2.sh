#!/bin/bash
# round-2 bench set on N GPUs of one box (run under gpurun [--gpus N]); JSON lines land in gpurun_out/r2_bench_<config>_<N>gpu.json
N=${1:-1}
run() {
  name=$1; shift
  if [ "$N" = "1" ]; then timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu --no-her "$@" > gpurun_out/r2_bench_${name}_1gpu.json 2>> gpurun_out/r2_bench_1gpu.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-her "$@" > gpurun_out/r2_bench_${name}_${N}gpu.json 2>> gpurun_out/r2_bench_${N}gpu.err; fi
}
run reach_joints
run reach_ee --control ee
run pick_and_place --task pick_and_place --control ee --envs 32768
if [ "$N" = "1" ]; then
  for t in push slide flip stack; do run $t --task $t --control ee; done
  run reach_joints_262144 --envs 262144
  run reach_ee_262144 --control ee --envs 262144
  run pick_and_place_131072 --task pick_and_place --control ee --envs 131072
  run reach_joints_dense --reward dense
fi
python scripts/show_ab.py gpurun_out/r2_bench_
