mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu > $O/pytest.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python tests/gpu_diag_sweep.py > $O/diag_sweep.txt 2>&1; echo "diag rc=$?"
tail -n 4 $O/pytest.txt | cut -c1-300
grep "device-resident\|profile ranges" $O/*.err | cut -c1-600
grep -A14 "DS2_BWD_LL = 0" $O/diag_sweep.txt | grep -A13 "trace fwd" | cut -c1-120
