mkdir -p gpurun_out/r2ab
O=gpurun_out/r2ab
timeout 1500 python -m pytest tests -q -x -m gpu > $O/pytest_all.txt 2>&1; echo "pytest_all rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --precision tf32 --no-cpu-baseline --no-stock-cuda > $O/bench_tf32.json 2> $O/bench_tf32.err; echo "bench tf32 rc=$?"
DS2_SANITIZE_BIG=1 timeout 1200 compute-sanitizer --tool memcheck --kernel-regex kns=ds2 --log-file $O/memcheck.txt python tests/gpu_sanitize_small.py > $O/memcheck_stdout.txt 2>&1; echo "memcheck rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --kernel-regex kns=ds2 --log-file $O/racecheck.txt python tests/gpu_sanitize_small.py > $O/racecheck_stdout.txt 2>&1; echo "racecheck rc=$?"
B="python bench.py --steps 2 --warmup 1 --no-stock-cuda --no-parity --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $B > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/aggregate_launches.py $O/launches.csv > $O/one_step.csv 2> $O/agg.err; echo "agg rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv2_wgrad_tc_kernel|conv1_wgrad_tc_kernel" -c 2 --csv --page raw --log-file $O/conv_wgrad_ncu_raw.csv $B > $O/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
for w in an4 unigru_lookahead stress; do
  timeout 900 python bench.py --steps 8 --warmup 3 --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"
done
tail -n 3 $O/pytest_all.txt; tail -n 2 $O/smoke.txt; grep "device-resident\|e2e:\|stock baseline done" $O/*.err | cut -c1-300; tail -n 3 $O/memcheck.txt; tail -n 3 $O/racecheck.txt; tail -n 5 $O/memcheck_stdout.txt; head -14 $O/one_step.csv
