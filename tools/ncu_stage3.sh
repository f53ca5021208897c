#!/bin/bash
# `ncu --set full` captures of the stage-3 GEMM shapes (M = 50176, C = 384 at batch 256): bash tools/ncu_stage3.sh <tag>
tag=${1:-ncu}
mkdir -p gpurun_out
cap() {
  timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/${tag}_$1" \
    python tools/prof_kernels.py "$4" 2 > "gpurun_out/${tag}_$1.log" 2>&1
  echo "$1 rc=$?"
}
cap s3_qkv gemm_tn_kernel 2 gemm_qkv
cap s3_dqkv gemm_tn_kernel 2 gemm_dqkv
cap s3_fc2 gemm_tn_kernel 2 gemm_fc2
cap s3_fc1 gemm_tn_kernel 2 gemm_fc1
