#!/bin/bash
cd /root/repo
L=gpurun_out/lab_run4.txt
: > $L
LAB_K2L_CTAS_PER_SM=2 timeout 300 profiles/lab/bin/lab_v2_ov1 ov1_k2l2cta 1250000 10 0 >> $L 2>&1
LAB_K2L_CTAS_PER_SM=1 timeout 300 profiles/lab/bin/lab_v2_ov1 ov1_k2l1cta 1250000 10 0 >> $L 2>&1
timeout 300 profiles/lab/bin/lab_v2_ov1_r72 ov1_r72 1250000 10 0 >> $L 2>&1
LAB_VAL_CTAS=296 timeout 300 profiles/lab/bin/lab_v2_ov1_r72 ov1_r72_296 1250000 10 0 >> $L 2>&1
cat $L
