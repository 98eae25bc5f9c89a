#!/bin/bash
cd /root/repo
L=gpurun_out/lab_run2.txt
: > $L
for b in lab_v2_ov1 lab_v2_ov2; do timeout 300 profiles/lab/bin/$b $b 1250000 10 0 >> $L 2>&1; done
cat $L
for b in lab_ref lab_v2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:range_lean -s 5 -c 1 -o gpurun_out/prof_$b -f profiles/lab/bin/$b $b 200000 3 0 > gpurun_out/ncu_$b.log 2>&1
  tail -3 gpurun_out/ncu_$b.log
done
