#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_kernels_gpu.py -k "talking" tests/test_dropout_gpu.py tests/test_parity_extra_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 900 python -m pytest tests/test_model_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
for f in 1 0; do
  SPE_TH_S16=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s$f.json 2> gpurun_out/bench_s$f.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_s$f.json').read().strip().splitlines()[-1])
    print("S16=$f value %.2f img/s  ms/step %.2f  e2e %.2f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
    for k, v in d['kernel_breakdown'].items(): print("  %-22s %8.3f ms/step  %6.1f launches" % (k, v['ms_per_step'], v['launches_per_step']))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_s$f.err').read()[-2000:])
PY
done
