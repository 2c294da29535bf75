#!/bin/bash
# usage: scripts/sweep_rj.sh <task> <control> <envs> <steps> "<groups list>" "<segments list>"   (under gpurun): PG_GROUPS x PG_SEGMENTS sweep, steady-state bench
T=$1; C=$2; N=$3; K=$4
for g in $5; do for s in $6; do
  PG_GROUPS=$g PG_SEGMENTS=$s python bench.py --task $T --control $C --envs $N --steps $K --warmup 5 --no-cpu --no-her 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$T $C groups=$g segments=$s: %.3e  %.3f ms'%(d['value'],d['ms_per_step']))"
done; done
