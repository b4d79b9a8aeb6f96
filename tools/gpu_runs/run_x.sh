mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "conv or train_step" > $O/pytest_conv.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
DS2_CONV_NO_CLUSTER=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda --no-parity > $O/bench_nocl.json 2> $O/bench_nocl.err; echo "bench rc=$?"
tail -n 5 $O/pytest_conv.txt
grep "device-resident\|profile ranges" $O/*.err | cut -c1-600
