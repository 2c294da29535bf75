for g in 4 8; do
 for t in "pick_and_place 32768" "push 65536" "stack 65536"; do set -- $t
  PG_GROUPS=$g timeout 200 python bench.py --task $1 --control ee --envs $2 --steps 50 --warmup 10 --no-cpu --no-her > gpurun_out/ab4_${1}_g$g.json 2> gpurun_out/ab4_err.log
 done
done
