#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
for f in 1 0; do
  SPE_GEMM_PDL=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p$f.json 2> gpurun_out/bench_p$f.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_p$f.json').read().strip().splitlines()[-1])
    print("PDL=$f value %.2f img/s  ms/step %.2f  e2e %.2f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
    print("   gemm %.3f ms/step" % d['kernel_breakdown']['gemm']['ms_per_step'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_p$f.err').read()[-2000:])
PY
done
