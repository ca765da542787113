#!/bin/bash
# ncu --set full of the round's changed training kernels: weight-gradient GEMM (8-warp vector-reduction epilogue) and the
# two-CTAs-per-SM plain GEMM instances, at C3 shapes through the block-level hooks
mkdir -p gpurun_out
for t in wgrad gemm2; do
  case $t in gemm2) k=gemm_tc;; wgrad) k=wgrad_tc;; esac
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -f -o gpurun_out/prof_${t}_r2 python tools/prof_kernels.py $t > gpurun_out/prof_$t.log 2>&1; echo "$t rc=$?"
done
ls -la gpurun_out/prof_wgrad_r2.ncu-rep gpurun_out/prof_gemm2_r2.ncu-rep
