mkdir -p gpurun_out/r2aa
O=gpurun_out/r2aa
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "conv or train_step or sinks or deferred" > $O/pytest_conv.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -n 3 $O/pytest_conv.txt | cut -c1-300
grep "device-resident\|profile ranges" $O/*.err | cut -c1-600
