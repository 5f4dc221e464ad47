#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|talking" -s 14 -c 14 -o gpurun_out/prof_attn -f python tools/prof_attn.py 2 > gpurun_out/ncu_attn.log 2>&1
tail -5 gpurun_out/ncu_attn.log
ls -la gpurun_out/*.ncu-rep
