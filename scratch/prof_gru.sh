#!/bin/bash
# fused GRU step (glin_tc3_kernel<T3_ACT_GRU>): plain decode first, then one --set full capture summarised on the box
python scratch/decode_only.py fp16x2 > gpurun_out/r2_gru_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_gru_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:glin_tc3_kernel --launch-skip 5 -c 1 -o gpurun_out/r2_gru_step python scratch/decode_only.py fp16x2 > gpurun_out/r2_gru_ncu.log 2>&1
python scratch/ncu_summary.py gpurun_out/r2_gru_step.ncu-rep > gpurun_out/r2_gru_step.summary.txt 2>&1
rm -f gpurun_out/r2_gru_step.ncu-rep
head -22 gpurun_out/r2_gru_step.summary.txt
