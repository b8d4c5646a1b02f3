OUT=gpurun_out/r02m; mkdir -p $OUT
( time timeout 1200 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ) 2> $OUT/bench.time; echo "bench rc=$?"; tail -3 $OUT/bench.err; cat $OUT/bench.time
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ) 2> $OUT/ref.time; echo "ref rc=$?"; tail -3 $OUT/bench_ref.err; cat $OUT/ref.time
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e',d['e2e']['value'],d['e2e']['single_caller']['value'],d['e2e']['compressed_points']['value'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','peak','frac','peak_per_clk_per_sm_wide','peak_per_clk_per_sm_pairs','whole_msm_frac')})
print('stage',d['stage_ms'])
print('sustained',d.get('sustained'))
bb=d['blindbid']
for k in ('prove','prove_large_batch','batch_verify','batch_verify_large','batch_verify_1024_total'):
    print(k,bb[k]['value'],bb[k].get('ms_per_batch'),bb[k].get('roofline',{}).get('frac'))
print(bb['batch_verify_1024_corrupted']); print(bb['prove_config3_sweep']); print(bb['single_request_ms'])
print('cpu',d.get('cpu_baseline'))
r=json.loads(open('gpurun_out/r02m/bench_ref.json').read().strip().splitlines()[-1])
print('ref',r['value'],r['blindbid'])
PY
