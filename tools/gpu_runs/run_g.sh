mkdir -p gpurun_out/r2g
B="python bench.py --steps 2 --warmup 1 --no-stock-cuda --no-parity --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2g/launches.csv $B > gpurun_out/r2g/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/aggregate_launches.py gpurun_out/r2g/launches.csv > gpurun_out/r2g/one_step.csv 2> gpurun_out/r2g/agg.err; echo "agg rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rnn_fwd_splitk|rnn_bwd_splitk|gemm_tc_kernel" -c 9 -o gpurun_out/r2g/r02_rnn_layer python tests/gpu_one_layer.py lstm fp16 > gpurun_out/r2g/ncu_layer.log 2>&1; echo "ncu layer rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv|ctc_alpha" -s 14 -c 8 -o gpurun_out/r2g/r02_conv_ctc python bench.py --steps 1 --warmup 1 --no-stock-cuda --no-parity --no-cpu-baseline > gpurun_out/r2g/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
for w in an4 unigru_lookahead stress; do
  timeout 900 python bench.py --steps 8 --warmup 3 --workload $w > gpurun_out/r2g/bench_$w.json 2> gpurun_out/r2g/bench_$w.err; echo "bench $w rc=$?"
done
head -30 gpurun_out/r2g/one_step.csv; grep "device-resident\|stock baseline done" gpurun_out/r2g/bench_*.err
