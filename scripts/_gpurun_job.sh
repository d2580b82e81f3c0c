mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) 2>&1 | tail -5
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
(time timeout 900 python bench.py) > gpurun_out/r02_bench_v6.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_v6.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['pageable_inputs'], d['roofline']['whole_step']['frac'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
for k,v in list(d['kernels'].items())[:8]: print(k, v['ms'], v['tflops'], v['gbs'])
for k in ('cloud_opt','cloud_opt_config5'):
    c=d.get(k); print(k, c['value'], c['roofline']['frac'], c.get('e2e',{}).get('value'), c.get('cpu_baseline',{}).get('value'))
PY
tail -3 gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none -k regex:align_stream -s 3 -c 1 -o gpurun_out/r02_prof_align_stream_config5 python scripts/ncu_target.py align50 > gpurun_out/ncu_align50.log 2>&1; tail -2 gpurun_out/ncu_align50.log
