#!/bin/bash
# usage: scripts/quick2.sh <tag>   (under gpurun): steady-state bench lines (60-step pre-roll) of the main configurations -> gpurun_out/q2_<tag>_*.json
TAG=$1
run() { python bench.py --task $1 --control $2 --envs $3 --steps $4 --warmup 5 --no-cpu --no-her > gpurun_out/q2_${TAG}_$1_$2.json 2> gpurun_out/q2_${TAG}_$1_$2.err; }
run reach joints 65536 60
run reach ee 65536 40
run pick_and_place ee 32768 30
run push ee 65536 20
run stack ee 65536 15
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/q2_${TAG}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split("q2_")[1], "%.3e env-steps/s  %.3f ms  e2e %.3e"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["episode_stats"]["success_rate"])
    except Exception as e: print(f,"ERR",e)
PY
