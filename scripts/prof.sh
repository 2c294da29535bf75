#!/bin/bash
# usage: scripts/prof.sh <task> <control> <envs> <skip step_kernel launches> <name>   (run under gpurun; writes gpurun_out/prof_<name>.*)
set -x
T=$1; C=$2; N=$3; SKIP=$4; NAME=$5
ARGS="--task $T --control $C --envs $N --steps 20 --warmup 5 --no-cpu --no-her"
timeout 300 python bench.py $ARGS > gpurun_out/prof_${NAME}_plain.json 2> gpurun_out/prof_${NAME}_plain.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $SKIP -c 1 -f -o gpurun_out/prof_$NAME python bench.py $ARGS > gpurun_out/prof_${NAME}_ncu.log 2>&1
ncu -i gpurun_out/prof_$NAME.ncu-rep --page raw --csv > gpurun_out/prof_${NAME}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$NAME.ncu-rep --page source --csv > gpurun_out/prof_${NAME}_src.csv 2>/dev/null
rm -f gpurun_out/prof_$NAME.ncu-rep
