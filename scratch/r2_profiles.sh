#!/bin/bash
# Round-2 profile captures (run on the GPU box).  Each capture: the program alone first (must exit 0), then under ncu; the report is
# summarised ON the box (scratch/ncu_summary.py) and deleted: only text travels back (gpurun_out/ is limited to 64 MiB).
export PROF_WARM=1 PROF_ITERS=1
cap() {   # cap <tag> <kernel regex> <count> <program...>
  tag=$1; rx=$2; cnt=$3; shift 3
  "$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "$tag: plain run failed"; tail -3 gpurun_out/${tag}_plain.log; return 1; }
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o gpurun_out/$tag "$@" > gpurun_out/${tag}_ncu.log 2>&1
  python scratch/ncu_summary.py gpurun_out/$tag.ncu-rep > gpurun_out/$tag.summary.txt 2>&1
  rm -f gpurun_out/$tag.ncu-rep
  grep -E "^## |gpu__time_duration|dram__bytes" gpurun_out/$tag.summary.txt | head -40
}
if [ "$1" != "launches-only" ]; then
cap r2_final_tc3 "glin_tc3" 4 python scratch/prof_kernels.py tc3,tc3raw,qkv,toout
cap r2_final_step_attn "reverse_step|node_attention_bulk" 2 python scratch/prof_kernels.py step,attn
cap r2_final_dense "node_attention_bulk|gru_sample|gru_head|sample_mix" 9 python scratch/prof_mix.py
# fused GRU step of the decoder (identity influence, fp16x2): the 6th launch of a decode
"$(dirname "$0")"/prof_gru.sh && cp gpurun_out/r2_gru_step.summary.txt gpurun_out/r2_final_gru_step.summary.txt
fi
export PROF_DENOISER=1
cap r2_final_attn_mix "node_attention_bulk_kernel<21, 1>|node_attention_bulk_kernelILi21ELb1" 1 python scratch/prof_mix.py
# launch list of the bench command (first 9000 launches: warm-up replays of the graph, the timed steps, the eager step)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2_final_launches_bench_plain.json 2> gpurun_out/r2_final_launches_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2_final_launches_ncu.log 2>&1
python scratch/launch_summary.py gpurun_out/r2_final_launches.csv > gpurun_out/r2_final_launches.summary.txt
head -30 gpurun_out/r2_final_launches.summary.txt
du -sh gpurun_out
