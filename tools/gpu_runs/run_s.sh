mkdir -p gpurun_out/r2s
DS2_FWD_SPLITK=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda --no-parity > gpurun_out/r2s/bench_fwd16.json 2> gpurun_out/r2s/bench_fwd16.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda --no-parity > gpurun_out/r2s/bench_default.json 2> gpurun_out/r2s/bench_default.err; echo "bench rc=$?"
grep "device-resident\|profile ranges" gpurun_out/r2s/*.err | cut -c1-500
