mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q -s -m gpu -k "fp16 or precision_16" > gpurun_out/r2f/pytest_fp16.txt 2>&1; echo "pytest_fp16 rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --precision fp16 > gpurun_out/r2f/bench_fp16.json 2> gpurun_out/r2f/bench_fp16.err; echo "bench rc=$?"
grep "fullsize\]" gpurun_out/r2f/pytest_fp16.txt; tail -n 4 gpurun_out/r2f/pytest_fp16.txt; grep "device-resident\|e2e\|stock baseline done\|profile ranges\|parity_fullsize" gpurun_out/r2f/bench_fp16.err | cut -c1-900
