#!/bin/bash
# usage: scripts/ab.sh <tag> [lib.so] [configs...]  -- 30 timed steps per configuration after the 60-step pre-roll (run under gpurun)
TAG=$1; LIB=$2; shift; shift
[ -n "$LIB" ] && export PANDA_B200_LIB=$PWD/$LIB
CONFIGS=${@:-"reach/joints/65536 reach/ee/65536 pick_and_place/ee/32768 push/ee/65536 stack/ee/65536"}
for c in $CONFIGS; do
  IFS=/ read t ctl n <<< "$c"
  timeout 300 python bench.py --task $t --control $ctl --envs $n --steps 30 --warmup 3 --no-cpu --no-her > gpurun_out/ab_${TAG}_${t}_${ctl}_${n}.json 2>> gpurun_out/ab_${TAG}_err.log
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${TAG}_${t}_${ctl}_${n}.json"))
    print("${TAG} $c: %.3e env-steps/s  %.3f ms/step  e2e %.3e  stats %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["episode_stats"]))
except Exception as e:
    print("${TAG} $c: FAILED", e)
PY
done
