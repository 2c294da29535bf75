set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/seg_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/seg_pytest.log
for s in 1 4 10 20; do
 for t in "pick_and_place ee 32768" "push ee 65536" "reach joints 65536" "stack ee 65536"; do
  set -- $t
  PG_SEGMENTS=$s timeout 200 python bench.py --task $1 --control $2 --envs $3 --steps 20 --warmup 5 --no-cpu --no-her > gpurun_out/seg_${1}_$s.json 2> gpurun_out/seg_err.log
 done
done
tail -3 gpurun_out/seg_pytest.log
