#!/bin/bash
# timing-only experiments on glin_tc3_kernel variants (results of the variants are intentionally wrong)
L=skeletondiffusion_b200/csrc
cp $L/libskeldiff_sm100a.so /tmp/lib_orig.so
for v in T3_EXP_HI_ONLY T3_EXP_ONE_MMA; do
  cp $L/libskeldiff_$v.so $L/libskeldiff_sm100a.so
  echo "variant $v"; (timeout 100 python scratch/prof_kernels.py tc3 2>&1 | tail -1)
done
cp /tmp/lib_orig.so $L/libskeldiff_sm100a.so
