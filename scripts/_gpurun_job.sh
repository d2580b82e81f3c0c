mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -4
timeout 900 python bench.py --skip-cpu-baseline --skip-cloud-opt > gpurun_out/r02_bench_v4.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_v4.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_step']['frac'])
for k,v in list(d['kernels'].items())[:8]: print(k, v['ms'], v['tflops'], v['gbs'])
PY
