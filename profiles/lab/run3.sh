#!/bin/bash
cd /root/repo
L=gpurun_out/lab_run3.txt
: > $L
LAB_VAL_CTAS=592 LAB_VAL_THREADS=32 timeout 300 profiles/lab/bin/lab_v2_ov1 ov1_592x32 1250000 10 0 >> $L 2>&1
LAB_VAL_CTAS=296 LAB_VAL_THREADS=64 timeout 300 profiles/lab/bin/lab_v2_ov1 ov1_296x64 1250000 10 0 >> $L 2>&1
LAB_VAL_CTAS=148 LAB_VAL_THREADS=128 timeout 300 profiles/lab/bin/lab_v2_ov1_r72 ov1_r72_148x128 1250000 10 0 >> $L 2>&1
LAB_VAL_CTAS=296 LAB_VAL_THREADS=128 timeout 300 profiles/lab/bin/lab_v2_ov1_r72 ov1_r72_296x128 1250000 10 0 >> $L 2>&1
LAB_VAL_CTAS=2368 LAB_VAL_THREADS=128 timeout 300 profiles/lab/bin/lab_v2_ov1 ov1_2368x128 1250000 10 0 >> $L 2>&1
timeout 300 profiles/lab/bin/lab_v2_mb4 v2_mb4 1250000 10 0 >> $L 2>&1
timeout 300 profiles/lab/bin/lab_ref_mb4 ref_mb4 1250000 10 0 >> $L 2>&1
cat $L
