mkdir -p gpurun_out/r2i
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r2i/pytest_all.txt 2>&1; echo "pytest_all rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --precision tf32 --no-cpu-baseline > gpurun_out/r2i/bench_tf32.json 2> gpurun_out/r2i/bench_tf32.err; echo "bench tf32 rc=$?"
tail -n 4 gpurun_out/r2i/pytest_all.txt; tail -n 2 gpurun_out/r2i/smoke.txt; grep "device-resident\|e2e:\|profile ranges\|stock baseline done" gpurun_out/r2i/*.err | cut -c1-500
