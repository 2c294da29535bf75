# usage: exp_watch.sh <tag>  -- one ncu capture of a mid-step Reach ee launch + per-warp timing, for A/B of solver variants
TAG=$1
ARGS="--task reach --control ee --envs 65536 --steps 20 --warmup 5 --no-cpu --no-her"
timeout 300 python bench.py $ARGS > gpurun_out/ew_${TAG}_plain.json 2> gpurun_out/ew_${TAG}.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 630 -c 1 -f -o gpurun_out/ew_$TAG python bench.py $ARGS > gpurun_out/ew_${TAG}_ncu.log 2>&1
ncu -i gpurun_out/ew_$TAG.ncu-rep --page raw --csv > gpurun_out/ew_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/ew_$TAG.ncu-rep --page source --csv > gpurun_out/ew_${TAG}_src.csv 2>/dev/null
rm -f gpurun_out/ew_$TAG.ncu-rep
PG_DEBUG_TIMING=1 python scripts/timing_debug.py reach ee 65536 > gpurun_out/ew_${TAG}_timing.log 2>&1
