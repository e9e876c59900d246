#!/bin/bash
# Runs on the GPU box (gpurun): the default bench (plain, then again under ncu for the launch list) and one
# `ncu --set full` capture per main kernel.  Everything lands in gpurun_out/; summaries are extracted here afterwards.
# usage: scratch/final_profiles.sh <tag>
TAG=${1:-r1_final}
O=gpurun_out
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || echo "bench failed"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
for k in tc3 tc ffma step attn; do
  case $k in
    tc3) rx=glin_tc3;; tc) rx=glin_tc_kernel;; ffma) rx=glin_gemm_f2;; step) rx=reverse_step;; attn) rx=node_attention;;
  esac
  timeout 120 python scratch/prof_kernels.py $k > $O/${TAG}_plain_$k.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o $O/${TAG}_$k -f \
      python scratch/prof_kernels.py $k > $O/${TAG}_ncu_$k.log 2>&1
  tail -1 $O/${TAG}_plain_$k.log
done
ls -la $O | grep $TAG
