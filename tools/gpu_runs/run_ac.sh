mkdir -p gpurun_out/r2ac
O=gpurun_out/r2ac
timeout 1200 python -m pytest tests -q -x -m gpu > $O/pytest_all.txt 2>&1; echo "pytest_all rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
B="python bench.py --steps 2 --warmup 1 --no-stock-cuda --no-parity --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $B > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/aggregate_launches.py $O/launches.csv > $O/one_step.csv 2> $O/agg.err; echo "agg rc=$?"
tail -n 3 $O/pytest_all.txt; tail -n 2 $O/smoke.txt; grep "device-resident\|e2e:\|stock baseline done\|profile ranges" $O/*.err | cut -c1-500
