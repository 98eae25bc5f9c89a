#!/bin/bash
cd /root/repo
L=gpurun_out/lab_run5.txt
: > $L
for j in ${JITS:-0}; do
  for b in $(ls profiles/lab/bin); do
    timeout 300 profiles/lab/bin/$b ${b}_j$j 1250000 10 0 $j >> $L 2>&1
  done
done
grep -v "offsets=" $L
