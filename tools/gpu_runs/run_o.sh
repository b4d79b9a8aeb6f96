mkdir -p gpurun_out/r2o
nvidia-smi -L > gpurun_out/r2o/gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -q -x -m gpu > gpurun_out/r2o/pytest_dist.txt 2>&1; echo "pytest dist rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2o/bench_n2.json 2> gpurun_out/r2o/bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-stock-cuda --no-cpu-baseline --no-parity > gpurun_out/r2o/bench_n1.json 2> gpurun_out/r2o/bench_n1.err; echo "bench n1 rc=$?"
tail -n 3 gpurun_out/r2o/pytest_dist.txt; grep "device-resident\|e2e:" gpurun_out/r2o/*.err; cat gpurun_out/r2o/bench_n2.json | cut -c1-300
