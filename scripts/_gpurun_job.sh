mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
(time timeout 900 python bench.py) > gpurun_out/r02_bench_v3.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_v3.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['whole_step']['frac'], d.get('cpu_baseline'))
for k in ('cloud_opt','cloud_opt_config5'):
    c=d.get(k); print(k, {kk: c[kk] for kk in c if kk not in ('config','roofline')} , c.get('roofline',{}).get('frac'))
PY
tail -5 gpurun_out/bench.err
