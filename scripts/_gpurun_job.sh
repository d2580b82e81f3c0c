mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/r02_bench_v2.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_v2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_step']['frac'])
for k,v in d['kernels'].items(): print(k, v['ms'], v['tflops'], v['gbs'])
print(json.dumps(d.get('cloud_opt'))[:900])
PY
tail -3 gpurun_out/bench.err
