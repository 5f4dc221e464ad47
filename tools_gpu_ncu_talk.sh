#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"talking_(fwd|bwd)_rows" -s 2 -c 2 -o gpurun_out/prof_talk -f python tools/prof_attn.py 2 > gpurun_out/ncu_talk.log 2>&1
tail -3 gpurun_out/ncu_talk.log
