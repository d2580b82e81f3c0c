mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_align_gpu.py -x -q 2>&1 | tail -2
(time timeout 900 python bench.py) > gpurun_out/r02_bench_v7.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_v7.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_step']['frac'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['share_of_step'])
for k,v in list(d['kernels'].items())[:9]: print(k, v['ms'], v['tflops'], v['gbs'])
for k in ('cloud_opt','cloud_opt_config5'):
    c=d.get(k); print(k, c['value'], c['roofline']['frac'], c['roofline']['traffic'], c.get('e2e',{}).get('value'))
PY
