#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05" -s 12 -c 4 -o gpurun_out/prof_gemm -f python tools/prof_attn.py 2 > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
