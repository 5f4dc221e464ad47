#!/bin/bash
mkdir -p gpurun_out
# kernels of gemm_micro in order: S fp32 (x23), S bf16 (x23), qkv, fc1gelu, fc2res ...: take one launch of each of the first two + PV-like via prof_attn
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05" -s 30 -c 1 -o gpurun_out/prof_gemm_sbf16 -f python tools/gemm_micro.py > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
