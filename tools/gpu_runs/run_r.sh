mkdir -p gpurun_out/r2r
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "conv_frontend or tf32 or precision_16 or full_model" > gpurun_out/r2r/pytest.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2r/bench.json 2> gpurun_out/r2r/bench.err; echo "bench rc=$?"
tail -n 5 gpurun_out/r2r/pytest.txt; grep "device-resident\|profile ranges\|parity_fullsize" gpurun_out/r2r/*.err | cut -c1-700
