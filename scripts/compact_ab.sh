set -x
make -C panda_lang_manip_b200/csrc -j8 > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/cmp_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cmp_pytest.log
for c in 0 1; do
 for t in "pick_and_place ee 32768" "push ee 65536" "reach ee 65536" "stack ee 65536"; do
  set -- $t
  PG_COMPACT=$c timeout 200 python bench.py --task $1 --control $2 --envs $3 --steps 20 --warmup 5 --no-cpu --no-her > gpurun_out/cmp_${1}_$c.json 2> gpurun_out/cmp_err_${1}_$c.log
 done
done
tail -3 gpurun_out/cmp_pytest.log
