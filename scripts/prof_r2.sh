#!/bin/bash
# round-2 profiling pass (run under gpurun AFTER the plain commands exited 0): launch list of the bench command, one ncu --set full capture of a
# steady-state step_kernel launch (Reach joints, PickAndPlace), captures of the HER kernels.  Outputs: gpurun_out/r2_*
set -x
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-her > gpurun_out/r2_plain_bench.json 2> gpurun_out/r2_plain_bench.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_reach_joints.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-her > gpurun_out/r2_launches.log 2>&1
for cfg in "reach joints 65536 62 r2_reach_joints" "pick_and_place ee 32768 1240 r2_pnp"; do
  set -- $cfg
  ARGS="--task $1 --control $2 --envs $3 --steps 20 --warmup 5 --no-cpu --no-her"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $4 -c 1 -f -o gpurun_out/prof_$5 python bench.py $ARGS > gpurun_out/prof_$5_ncu.log 2>&1
  ncu -i gpurun_out/prof_$5.ncu-rep --page raw --csv > gpurun_out/prof_$5_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$5.ncu-rep
done
timeout 200 python scripts/her_prof.py > gpurun_out/r2_her_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none -k regex:"reward_kernel|her_relabel_kernel" -f -o gpurun_out/prof_r2_her python scripts/her_prof.py > gpurun_out/prof_r2_her_ncu.log 2>&1
ncu -i gpurun_out/prof_r2_her.ncu-rep --page raw --csv > gpurun_out/prof_r2_her_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r2_her.ncu-rep
