set -x
for s in 4 10 20; do for g in 1 2; do
  PG_SEGMENTS=$s PG_GROUPS=$g timeout 200 python bench.py --task reach --control ee --envs 65536 --steps 20 --warmup 5 --no-cpu --no-her > gpurun_out/ab2_reach_ee_s${s}_g$g.json 2> gpurun_out/ab2_err.log
done; done
for n in 65536 131072; do
  PG_GROUPS=4 timeout 300 python bench.py --task pick_and_place --control ee --envs $n --steps 20 --warmup 5 --no-cpu --no-her > gpurun_out/ab2_pnp_n${n}.json 2> gpurun_out/ab2_err.log
done
