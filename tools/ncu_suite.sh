#!/bin/bash
# One `ncu --set full` capture per hot kernel at the bench's stage-1 shapes (batch 256); reports land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/ncu_suite.sh r01e'
tag=${1:-ncu}
mkdir -p gpurun_out
cap() {   # cap <name> <kernel regex> <skip> <what> [stage]
  timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/${tag}_$1" \
    python tools/prof_kernels.py "$4" ${5:-0} > "gpurun_out/${tag}_$1.log" 2>&1
  echo "$1 rc=$?"
}
cap gemm_fc1 gemm_tn_kernel 2 gemm_fc1
cap gemm_outproj gemm_tn_kernel 2 gemm_outproj
cap gemm_qkv gemm_tn_kernel 2 gemm_qkv
cap gemm_wgrad_small gemm_tn_kernel 2 gemm_wgrad_small
cap attn_bwd window_attn_bwd 2 attn_bwd
cap attn_fwd window_attn_fwd 2 attn_fwd
cap ln_fwd layernorm_fwd 2 ln
cap ln_bwd layernorm_bwd 2 ln
cap gallery_filter cosine_filter 3 gallery_bench
cap gallery_rerank rerank_kernel 1 gallery_bench
ls -la gpurun_out/
