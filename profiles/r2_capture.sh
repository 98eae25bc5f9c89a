#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU): launch list of one bench step per config, and --set full captures of
# the kernels the bench line names.  Times under ncu are serialised / cold; bench.py's CUDA-event numbers are the
# reported ones.  Reports are summarised on the box (gpurun brings back at most 64 MiB); only the dominant kernel's
# report travels.
cd /root/repo
R=/tmp/r2rep; mkdir -p $R gpurun_out
B="python bench.py --steps 1 --warmup 1 --series-per-gpu 200000 --hist-per-gpu 20000 --wide-rows-per-gpu 2000000 --groups 16000 --e2e-series 0 --no-cpu-baseline --jitter-variant-ms 0"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.log 2>&1
# range_lean_kernel<FN, FLAGS, GROUPED, UNI>: every call launches the general and the uniform-cadence variant and the one
# cadence_probe_kernel does not pick returns at once.  Headline data (scrapes on the schedule): the uniform variants run.
N="--set full --clock-control none --kernel-name-base demangled"
ncu $N --import-source on -k regex:'lean_kernel.*bool.0, .bool.0, .bool.1>' -s 1 -c 1 -f -o $R/prof_r2_lean $B > gpurun_out/r2_ncu_lean.log 2>&1
ncu $N -k regex:'lean_kernel.*bool.0, .bool.1, .bool.1>' -s 1 -c 1 -f -o $R/prof_r2_grouped $B > gpurun_out/r2_ncu_grouped.log 2>&1
ncu $N -k regex:'histogram_fold|series_offsets|column_reduce_stage1|histogram_uniform|cadence_probe' -s 8 -c 8 -f -o $R/prof_r2_other $B > gpurun_out/r2_ncu_other.log 2>&1
# the jittered variant of the generator: the general first tier
ncu $N -k regex:'lean_kernel.*bool.0, .bool.0, .bool.0>' -s 1 -c 1 -f -o $R/prof_r2_general $B --workload rate --jitter-ms 1000 > gpurun_out/r2_ncu_general.log 2>&1
ls -la $R
# 200000 series x 1000 samples per launch
python profiles/summarize_ncu.py --traffic-json gpurun_out/r2_traffic.json 2e8 $R/prof_r2_lean.ncu-rep $R/prof_r2_general.ncu-rep $R/prof_r2_grouped.ncu-rep $R/prof_r2_other.ncu-rep > gpurun_out/r2_ncu_summary.md
ncu -i $R/prof_r2_lean.ncu-rep --page source --csv > gpurun_out/r2_lean_source.csv 2>/dev/null
ncu -i $R/prof_r2_lean.ncu-rep --page details > gpurun_out/r2_lean_details.txt 2>/dev/null
ncu -i $R/prof_r2_general.ncu-rep --page details > gpurun_out/r2_general_details.txt 2>/dev/null
ncu -i $R/prof_r2_grouped.ncu-rep --page details > gpurun_out/r2_grouped_details.txt 2>/dev/null
cp $R/prof_r2_lean.ncu-rep gpurun_out/
