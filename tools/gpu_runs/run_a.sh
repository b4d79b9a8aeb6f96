mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/gpu.txt
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_dist.py -q -s -m gpu > gpurun_out/r2a/pytest_new.txt 2>&1; echo "pytest_new rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
timeout 600 compute-sanitizer --tool memcheck --kernel-regex kns=ds2 --log-file gpurun_out/r2a/memcheck.txt python tests/gpu_sanitize_small.py > gpurun_out/r2a/memcheck_stdout.txt 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=ds2 --log-file gpurun_out/r2a/racecheck.txt python tests/gpu_sanitize_small.py > gpurun_out/r2a/racecheck_stdout.txt 2>&1; echo "racecheck rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_dist.py > gpurun_out/r2a/pytest_old.txt 2>&1; echo "pytest_old rc=$?"
tail -3 gpurun_out/r2a/pytest_new.txt gpurun_out/r2a/smoke.txt gpurun_out/r2a/pytest_old.txt
