#!/bin/bash
# usage: scripts/ab_trees.sh <tag> <tree>...   -- the same short bench lines from several checked-out + built trees (build_ab/<commit>, "." = this tree)
TAG=$1; shift
ROOT=$(pwd)
for t in "$@"; do
  name=$(echo $t | tr '/.' '__')
  cd $ROOT/$t
  for cfg in "--task reach --control joints" "--task reach --control ee" "--task pick_and_place --control ee --envs 32768" "--task push --control ee"; do
    c=$(echo $cfg | tr -d ' -' )
    timeout 300 python bench.py $cfg --steps 30 --warmup 5 --no-cpu --no-her > $ROOT/gpurun_out/ab_${TAG}_${name}_$c.json 2>> $ROOT/gpurun_out/ab_${TAG}.err
  done
done
cd $ROOT
python scripts/show_ab.py gpurun_out/ab_${TAG}_
