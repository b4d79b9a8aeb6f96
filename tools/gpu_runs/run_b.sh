mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "tf32" > gpurun_out/r2b/pytest_tf32.txt 2>&1; echo "pytest_tf32 rc=$?"
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_dist.py -q -s -m gpu > gpurun_out/r2b/pytest_new.txt 2>&1; echo "pytest_new rc=$?"
timeout 600 python tests/gpu_diag_sweep.py > gpurun_out/r2b/diag_sweep.txt 2>&1; echo "diag rc=$?"
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err; echo "bench rc=$?"
DS2_FWD_LL=0 DS2_BWD_LL=0 timeout 900 python bench.py --steps 8 --warmup 3 --no-stock-cuda --no-cpu-baseline --no-parity > gpurun_out/r2b/bench_noll.json 2> gpurun_out/r2b/bench_noll.err; echo "bench_noll rc=$?"
tail -n 4 gpurun_out/r2b/pytest_tf32.txt; tail -n 8 gpurun_out/r2b/pytest_new.txt; grep "device-resident" gpurun_out/r2b/*.err
