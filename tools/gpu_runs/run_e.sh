mkdir -p gpurun_out/r2e
./tools/cluster_query > gpurun_out/r2e/cluster_query.txt 2>&1
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r2e/pytest_all.txt 2>&1; echo "pytest_all rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err; echo "bench rc=$?"
cat gpurun_out/r2e/cluster_query.txt; tail -n 6 gpurun_out/r2e/pytest_all.txt; tail -n 2 gpurun_out/r2e/smoke.txt; grep "device-resident\|e2e\|stock baseline done" gpurun_out/r2e/bench.err
