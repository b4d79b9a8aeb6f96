mkdir -p gpurun_out/r2n
for i in 1 2; do
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "tf32 or fp16 or fullsize or full_size or repeat or spiky or deferred" > gpurun_out/r2n/pytest_$i.txt 2>&1; echo "rc=$?"
grep -n "AssertionError\|passed\|failed" gpurun_out/r2n/pytest_$i.txt | cut -c1-600
done
