#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU): launch list of one bench step per config, and --set full captures of
# the kernels the bench line names.  Times under ncu are serialised / cold; bench.py's CUDA-event numbers are the
# reported ones.
cd /root/repo
B="python bench.py --steps 1 --warmup 1 --series-per-gpu 200000 --hist-per-gpu 20000 --wide-rows-per-gpu 2000000 --groups 16000 --e2e-series 0 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'range_lean|histogram_quantile|series_offsets|column_reduce_stage1' -c 14 -f -o gpurun_out/prof_r2 $B > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
