set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest_gpu_v1.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_v1.log
timeout 300 python scripts/align_stream_timeline.py 8 > gpurun_out/r02_align_stream_timeline_v1.jsonl 2> gpurun_out/tl.err; cat gpurun_out/r02_align_stream_timeline_v1.jsonl; tail -3 gpurun_out/tl.err
timeout 600 python scripts/align_config5.py c3 n24 c5 > gpurun_out/r02_align_config5_v3.jsonl 2> gpurun_out/c5.err; cat gpurun_out/r02_align_config5_v3.jsonl; tail -3 gpurun_out/c5.err
(time timeout 900 python bench.py) > gpurun_out/r02_bench_v1.json 2> gpurun_out/bench.err; cat gpurun_out/r02_bench_v1.json; tail -5 gpurun_out/bench.err
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 3) > gpurun_out/r02_bench_ref_v1.json 2> gpurun_out/benchref.err; cat gpurun_out/r02_bench_ref_v1.json; tail -5 gpurun_out/benchref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_stream -s 3 -c 1 -o gpurun_out/r02_prof_align_stream_v3 python scripts/ncu_target.py align > gpurun_out/ncu_align.log 2>&1; tail -3 gpurun_out/ncu_align.log
