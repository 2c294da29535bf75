TAG=$1
timeout 300 python bench.py --no-cpu --no-her > gpurun_out/q_${TAG}_reach_joints.json 2> gpurun_out/q_err.log
timeout 300 python bench.py --control ee --no-cpu --no-her > gpurun_out/q_${TAG}_reach_ee.json 2>> gpurun_out/q_err.log
timeout 300 python bench.py --task pick_and_place --control ee --envs 32768 --no-cpu --no-her > gpurun_out/q_${TAG}_pnp.json 2>> gpurun_out/q_err.log
timeout 300 python bench.py --task push --control ee --no-cpu --no-her > gpurun_out/q_${TAG}_push.json 2>> gpurun_out/q_err.log
timeout 300 python bench.py --task stack --control ee --no-cpu --no-her > gpurun_out/q_${TAG}_stack.json 2>> gpurun_out/q_err.log
