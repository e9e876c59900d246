#!/bin/bash
# round-2 profile captures (run on the GPU box): plain run first, then ncu; reports small enough to travel back (< 64 MiB in total)
set -x
export PROF_WARM=1 PROF_ITERS=1
python scratch/prof_kernels.py tc3,tc3raw,qkv,toout,step,attn > gpurun_out/r2_prof_identity_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"glin_tc3|reverse_step|node_attention_bulk" -c 12 -o gpurun_out/r2_final_identity python scratch/prof_kernels.py tc3,tc3raw,qkv,toout,step,attn > gpurun_out/r2_prof_identity_ncu.log 2>&1
python scratch/prof_mix.py > gpurun_out/r2_prof_dense_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"node_attention_bulk|gru_sample|gru_head|sample_mix" -c 14 -o gpurun_out/r2_final_dense python scratch/prof_mix.py > gpurun_out/r2_prof_dense_ncu.log 2>&1
ls -la gpurun_out/
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_launches_bench_plain.json 2> gpurun_out/r2_final_launches_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_launches_ncu.log 2>&1
ls -la gpurun_out/
