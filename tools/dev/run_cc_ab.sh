#!/bin/bash
cd /root/repo
for f in 1 ""; do
  SPE_CACHE_CONFIG=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_c.json').read().strip().splitlines()[-1])
    print("CACHECFG='$f' value %.2f img/s  ms/step %.2f  e2e %.2f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
    print("   gemm %.3f ms/step  layernorm %.3f" % (d['kernel_breakdown']['gemm']['ms_per_step'], d['kernel_breakdown']['layernorm']['ms_per_step']))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_c.err').read()[-2000:])
PY
done
