#!/bin/bash
# launch list (ncu per-launch durations) of the dense-influence kernels: scratch/prof_mix.py with one Denoiser forward
export PROF_ITERS=2 PROF_DENOISER=1
python scratch/prof_mix.py > gpurun_out/r2_dense_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_dense_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_dense_launches.csv python scratch/prof_mix.py > gpurun_out/r2_dense_ncu.log 2>&1
python scratch/launch_summary.py gpurun_out/r2_dense_launches.csv > gpurun_out/r2_dense_launches.summary.txt
head -14 gpurun_out/r2_dense_launches.summary.txt
