set -x
timeout 600 python -m pytest tests/test_align_gpu.py -x -q 2>&1 | tail -4
timeout 600 python scripts/align_config5.py c3 n24 c5 > gpurun_out/r02_align_config5_v3.jsonl 2> gpurun_out/r02_align_config5_v3.err
cat gpurun_out/r02_align_config5_v3.jsonl; tail -5 gpurun_out/r02_align_config5_v3.err
ncu --set full --clock-control none --import-source on -k regex:align_stream -s 3 -c 1 -o gpurun_out/r02_prof_align_stream_v3 python scripts/ncu_target.py align > gpurun_out/ncu_align.log 2>&1
