set -x
bash scripts/prof.sh reach joints 65536 62 r1_reach_joints
bash scripts/prof.sh pick_and_place ee 32768 1240 r1_pnp
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_reach_joints.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-her > gpurun_out/launches_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1_pnp.csv python bench.py --task pick_and_place --control ee --envs 32768 --steps 2 --warmup 1 --no-cpu --no-her > gpurun_out/launches_ncu2.log 2>&1
PG_DEBUG_TIMING=1 python scripts/timing_debug.py pick_and_place ee 32768 > gpurun_out/timing_pnp2.log 2>&1
