mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu_v2.json 2> gpurun_out/b2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu_v2.json'))
print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/config5_pipeline.py 50 2>/dev/null | tail -1
