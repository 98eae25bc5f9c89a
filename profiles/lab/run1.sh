#!/bin/bash
# gpurun payload: lab variants (1.25 M series), then the GPU test-suite on the rebuilt library
cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/lab_run1.txt
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L
for b in lab_ref lab_v2 lab_v2_mb2 lab_v2_ov1 lab_v2_ov2; do
  timeout 300 profiles/lab/bin/$b $b 1250000 10 0 >> $L 2>&1
done
# reset-variant data: plain variant hands everything off (timing is of the hand-off only); FLAGS variant keeps them
for b in lab_ref_flags lab_v2_flags; do
  timeout 300 profiles/lab/bin/$b ${b}_resets 1250000 10 1 >> $L 2>&1
  timeout 300 profiles/lab/bin/$b ${b}_noresets 1250000 10 0 >> $L 2>&1
done
# irregular data (jitter up to a full scrape interval): more missed guesses
for b in lab_ref lab_v2; do
  timeout 300 profiles/lab/bin/$b ${b}_jit14999 1250000 10 0 14999 >> $L 2>&1
done
cat $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gputest_run1.txt
