mkdir -p gpurun_out/r2h
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -m gpu > gpurun_out/r2h/pytest_parity.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-stock-cuda --no-cpu-baseline > gpurun_out/r2h/bench_defer.json 2> gpurun_out/r2h/bench_defer.err; echo "bench defer rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-stock-cuda --no-cpu-baseline --no-parity --no-defer > gpurun_out/r2h/bench_nodefer.json 2> gpurun_out/r2h/bench_nodefer.err; echo "bench nodefer rc=$?"
tail -n 5 gpurun_out/r2h/pytest_parity.txt; grep "device-resident\|e2e:\|profile ranges" gpurun_out/r2h/*.err | cut -c1-600
