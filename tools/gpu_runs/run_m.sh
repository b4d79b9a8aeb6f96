mkdir -p gpurun_out/r2m
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "tf32 or fp16 or fullsize or full_size or repeat or spiky or deferred" > gpurun_out/r2m/pytest.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2m/bench.json 2> gpurun_out/r2m/bench.err; echo "bench rc=$?"
DS2_FWD_LL=0 timeout 600 python tests/gpu_diag_sweep.py > gpurun_out/r2m/diag_sweep.txt 2>&1; echo "diag rc=$?"
tail -n 3 gpurun_out/r2m/pytest.txt; grep "device-resident\|profile ranges" gpurun_out/r2m/*.err | cut -c1-420; grep -A13 "trace fwd\|trace bwd" gpurun_out/r2m/diag_sweep.txt | tail -32
