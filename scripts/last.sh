timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_last.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_last.log
timeout 60 python bench.py --no-cpu --no-her > gpurun_out/q_last_reach_joints.json 2> gpurun_out/q_last.err
timeout 60 python bench.py --control ee --no-cpu --no-her > gpurun_out/q_last_reach_ee.json 2>> gpurun_out/q_last.err
timeout 60 python bench.py --task pick_and_place --control ee --envs 32768 --no-cpu --no-her > gpurun_out/q_last_pnp.json 2>> gpurun_out/q_last.err
