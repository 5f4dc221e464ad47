#!/bin/bash
# first GPU contact: kernel-level parity, one pytest process per group so a trap in one does not poison the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for k in test_gemm_majors test_gemm_epilogue test_gemm_batched test_linear_fn test_ffn_fn test_layernorm_fn test_attention_fn test_talking test_patch_embed test_match_cost test_lsap_kernel test_criterion_kernels; do
  echo "=== $k" >> gpurun_out/first.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "$k" -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/first.log
done
tail -150 gpurun_out/first.log
