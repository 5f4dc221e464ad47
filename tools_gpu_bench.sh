#!/bin/bash
mkdir -p gpurun_out
./tools_gpu_run.sh third "tests/test_model_gpu.py -k tiny" "tests/test_model_gpu.py -k cfg1" > /dev/null 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "--- third"; tail -30 gpurun_out/third.log | cut -c1-300
echo "--- smoke"; tail -5 gpurun_out/smoke.log
echo "--- bench"; tail -c 6000 gpurun_out/bench1.json; tail -20 gpurun_out/bench1.err
