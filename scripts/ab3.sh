for s in 4 10 20; do for g in 1 2 4; do
  PG_SEGMENTS=$s PG_GROUPS=$g timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu --no-her > gpurun_out/ab3_rj_s${s}_g$g.json 2> gpurun_out/ab3_err.log
done; done
for g in 2 8; do
  PG_GROUPS=$g timeout 200 python bench.py --control ee --steps 50 --warmup 10 --no-cpu --no-her > gpurun_out/ab3_re_s20_g$g.json 2> gpurun_out/ab3_err.log
done
