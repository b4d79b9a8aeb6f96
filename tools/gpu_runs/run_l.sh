mkdir -p gpurun_out/r2l
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu > gpurun_out/r2l/pytest.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-stock-cuda --workload stress > gpurun_out/r2l/bench_stress.json 2> gpurun_out/r2l/bench_stress.err; echo "bench stress rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda --workload unigru_lookahead > gpurun_out/r2l/bench_unigru.json 2> gpurun_out/r2l/bench_unigru.err; echo "bench unigru rc=$?"
tail -n 3 gpurun_out/r2l/pytest.txt; grep "device-resident\|profile ranges" gpurun_out/r2l/*.err | cut -c1-420
