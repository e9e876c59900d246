#!/bin/bash
# Per-source-line stall profile of the glin_tc3 epilogue variants (waits expanded at their call sites: MBAR_WAIT_AT)
export PROF_WARM=1 PROF_ITERS=1
python scratch/prof_kernels.py tc3,tc3raw > gpurun_out/r2_epi_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_epi_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:glin_tc3_kernel --launch-skip 1 --launch-count 3 -o gpurun_out/r2_epi python scratch/prof_kernels.py tc3,tc3raw > gpurun_out/r2_epi_ncu.log 2>&1
python scratch/ncu_summary.py gpurun_out/r2_epi.ncu-rep > gpurun_out/r2_epi.summary.txt 2>&1
rm -f gpurun_out/r2_epi.ncu-rep
grep -E "^## |gpu__time_duration" gpurun_out/r2_epi.summary.txt | head -20
