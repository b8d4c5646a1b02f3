OUT=gpurun_out/r02x; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
python bench.py --no-cpu --sustained-s 0.5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02x/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], d['e2e']['compressed_points']['value'])
bb=d['blindbid']
for k in ('prove','prove_large_batch','batch_verify','batch_verify_large'): print(k, round(bb[k]['value']), round(bb[k]['ms_per_batch'],2))
print(bb['single_request_ms'])
PY
