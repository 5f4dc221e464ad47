#!/bin/bash
# round-end evidence: launch list of one step, ncu --set full on the hot kernels, default bench, reference arm
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/step_once.py > gpurun_out/launchlist.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|talking|softmax_(fwd|bwd)_warp|layernorm" -s 20 -c 24 -o gpurun_out/prof_hot -f python tools/prof_attn.py 2 > gpurun_out/ncu_hot.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 1500 gpurun_out/bench_default.json; echo; cat gpurun_out/bench_reference.json | cut -c1-600; wc -l gpurun_out/launches.csv; ls -la gpurun_out/prof_hot.ncu-rep
