mkdir -p gpurun_out/r2q
timeout 600 python tests/gpu_diag_gemm16.py > gpurun_out/r2q/gemm16.txt 2>&1; echo "gemm16 rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "gemm" > gpurun_out/r2q/pytest_gemm.txt 2>&1; echo "pytest rc=$?"
cat gpurun_out/r2q/gemm16.txt; tail -n 3 gpurun_out/r2q/pytest_gemm.txt
