bash scripts/prof.sh reach joints 65536 62 r1_reach_joints
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_reach_joints.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-her > gpurun_out/launches_ncu.log 2>&1
