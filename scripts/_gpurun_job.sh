mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_stream -s 3 -c 1 -o gpurun_out/r02_prof_align_stream_v4 python scripts/ncu_target.py align > gpurun_out/ncu_align.log 2>&1; tail -2 gpurun_out/ncu_align.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 2 -c 1 -o gpurun_out/r02_prof_attn_tc3_v3 python scripts/ncu_attn_target.py 3 > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_ncu_launches_v1.csv python scripts/ncu_target.py forward 32 > gpurun_out/ncu_fwd.log 2>&1; tail -2 gpurun_out/ncu_fwd.log; wc -l gpurun_out/r02_ncu_launches_v1.csv
