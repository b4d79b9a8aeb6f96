mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "conv or train_step or sinks" > $O/pytest_conv.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -n 12 $O/pytest_conv.txt | cut -c1-300
grep "device-resident\|profile ranges" $O/*.err | cut -c1-600
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y/bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['parity_fullsize']['grad_rel_l2'])[:200]); print(d['parity_fullsize']['logits_rel'])
PY
