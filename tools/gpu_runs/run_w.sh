mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu > $O/pytest_parity.txt 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -n 3 $O/pytest_parity.txt
grep "device-resident\|profile ranges" $O/*.err | cut -c1-600
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w/bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['parity_fullsize']['grad_rel_l2'])[:300]); print(d['parity_fullsize']['logits_rel'])
PY
