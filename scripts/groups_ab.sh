set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/grp_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/grp_pytest.log
for c in 1 2 4 8; do
 for t in "pick_and_place ee 32768" "push ee 65536" "reach ee 65536" "reach joints 65536" "stack ee 65536"; do
  set -- $t
  PG_GROUPS=$c timeout 200 python bench.py --task $1 --control $2 --envs $3 --steps 20 --warmup 5 --no-cpu --no-her > gpurun_out/grp_${1}_${2}_$c.json 2> gpurun_out/grp_err_${1}_$c.log
 done
done
tail -3 gpurun_out/grp_pytest.log
