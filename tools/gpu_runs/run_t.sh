mkdir -p gpurun_out/r2t
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "conv or optim or adam or sgd or train_step" > gpurun_out/r2t/pytest_conv.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2t/bench.json 2> gpurun_out/r2t/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2t/pytest_conv.txt
grep "device-resident\|profile ranges\|parity" gpurun_out/r2t/*.err | cut -c1-600
