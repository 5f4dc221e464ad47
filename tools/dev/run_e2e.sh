#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print("value %.2f img/s  ms/step %.3f  e2e %.2f (%.3f ms)" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
PY
tail -3 gpurun_out/bench_e2e.err
