# full single-GPU bench set (run under gpurun); JSON lines land in gpurun_out/bench_<tag>.json
set -x
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}_reach_joints.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --control ee --no-cpu --no-her > gpurun_out/bench_${TAG}_reach_ee.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --task pick_and_place --control ee --envs 32768 --no-her > gpurun_out/bench_${TAG}_pnp.json 2>> gpurun_out/bench_${TAG}.err
for t in push slide stack flip; do
  timeout 300 python bench.py --task $t --control ee --no-cpu --no-her > gpurun_out/bench_${TAG}_$t.json 2>> gpurun_out/bench_${TAG}.err
done
timeout 300 python bench.py --envs 262144 --no-cpu --no-her > gpurun_out/bench_${TAG}_reach_joints_262144.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --task pick_and_place --control ee --envs 131072 --no-cpu --no-her > gpurun_out/bench_${TAG}_pnp_131072.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --control ee --envs 262144 --no-cpu --no-her > gpurun_out/bench_${TAG}_reach_ee_262144.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --reward dense --no-cpu --no-her > gpurun_out/bench_${TAG}_reach_joints_dense.json 2>> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/pytest_gpu_$TAG.log
