mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/config5_pipeline.py 50 > gpurun_out/r02_config5_pipeline_2gpu.jsonl 2> gpurun_out/c5p.err; cat gpurun_out/r02_config5_pipeline_2gpu.jsonl; tail -5 gpurun_out/c5p.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu_v1.json 2> gpurun_out/b2.err; head -c 700 gpurun_out/r02_bench_2gpu_v1.json; tail -3 gpurun_out/b2.err
