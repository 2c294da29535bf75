for s in 2 5; do PG_SEGMENTS=$s timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu --no-her > gpurun_out/ab5_rj_s$s.json 2> gpurun_out/ab5_err.log; done
PG_SEGMENTS=10 timeout 200 python bench.py --task pick_and_place --control ee --envs 32768 --steps 50 --warmup 10 --no-cpu --no-her > gpurun_out/ab5_pnp_s10.json 2>> gpurun_out/ab5_err.log
PG_SEGMENTS=10 timeout 200 python bench.py --control ee --steps 50 --warmup 10 --no-cpu --no-her > gpurun_out/ab5_re_s10.json 2>> gpurun_out/ab5_err.log
