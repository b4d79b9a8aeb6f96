mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r2p/pytest_all.txt 2>&1; echo "pytest_all rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --precision tf32 --no-cpu-baseline > gpurun_out/r2p/bench_tf32.json 2> gpurun_out/r2p/bench_tf32.err; echo "bench tf32 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p/bench_ref.json 2> gpurun_out/r2p/bench_ref.err; echo "bench ref rc=$?"
B="python bench.py --steps 2 --warmup 1 --no-stock-cuda --no-parity --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2p/launches.csv $B > gpurun_out/r2p/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/aggregate_launches.py gpurun_out/r2p/launches.csv > gpurun_out/r2p/one_step.csv 2> gpurun_out/r2p/agg.err; echo "agg rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rnn_fwd_splitk|rnn_bwd_splitk" -c 2 -o gpurun_out/r2p/r02_sweeps_final python tests/gpu_one_layer.py lstm fp16 > gpurun_out/r2p/ncu_layer.log 2>&1; echo "ncu layer rc=$?"
timeout 600 python tests/gpu_diag_sweep.py > gpurun_out/r2p/diag_sweep.txt 2>&1; echo "diag rc=$?"
for w in an4 unigru_lookahead stress; do
  timeout 900 python bench.py --steps 8 --warmup 3 --workload $w > gpurun_out/r2p/bench_$w.json 2> gpurun_out/r2p/bench_$w.err; echo "bench $w rc=$?"
done
tail -n 3 gpurun_out/r2p/pytest_all.txt; tail -n 2 gpurun_out/r2p/smoke.txt; grep "device-resident\|e2e:\|stock baseline done" gpurun_out/r2p/*.err | cut -c1-300; head -12 gpurun_out/r2p/one_step.csv
