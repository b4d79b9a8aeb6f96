mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err; echo "bench ref n8 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-stock-cuda --no-parity --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"
grep "device-resident\|e2e:" $O/*.err | cut -c1-300; tail -c 600 $O/bench_n8.json; tail -c 400 $O/bench_ref_n8.json
