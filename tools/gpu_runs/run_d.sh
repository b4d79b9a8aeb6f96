mkdir -p gpurun_out/r2d
timeout 300 ./tools/ll_microbench > gpurun_out/r2d/ll_microbench.txt 2>&1; echo "microbench rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -m gpu -k "sinks or direct" > gpurun_out/r2d/pytest_sinks.txt 2>&1; echo "pytest rc=$?"
cat gpurun_out/r2d/ll_microbench.txt; tail -n 5 gpurun_out/r2d/pytest_sinks.txt
