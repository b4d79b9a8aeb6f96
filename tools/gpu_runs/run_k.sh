mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu > gpurun_out/r2k/pytest.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench.err; echo "bench rc=$?"
timeout 600 python tests/gpu_diag_sweep.py > gpurun_out/r2k/diag_sweep.txt 2>&1; echo "diag rc=$?"
tail -n 4 gpurun_out/r2k/pytest.txt; grep "device-resident\|e2e:\|profile ranges\|stock baseline done" gpurun_out/r2k/*.err | cut -c1-500; sed -n 1,40p gpurun_out/r2k/diag_sweep.txt
