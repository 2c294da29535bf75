#!/bin/bash
# usage: scripts/launch_list.sh <name> <task> <ctrl> <envs> <launches per step> -- ncu per-launch durations of 2 steady-state steps (after 62 steps)
NAME=$1; T=$2; C=$3; N=$4; LPS=$5
SKIP=$((3 + 62 * LPS)); CNT=$((2 * LPS))
timeout 600 ncu --metrics gpu__time_duration.sum,launch__shared_mem_per_block_dynamic,launch__grid_size --clock-control none --launch-skip $SKIP --launch-count $CNT --csv --log-file gpurun_out/ll_$NAME.csv python scripts/steps.py $T $C $N 62 3 > gpurun_out/ll_$NAME.log 2>&1
