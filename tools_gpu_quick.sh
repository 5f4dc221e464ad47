#!/bin/bash
# usage: tools_gpu_quick.sh "<pytest args>" [bench args]   -- one pytest process + optional short bench
mkdir -p gpurun_out
eval "timeout 1200 python -m pytest $1 -q -x -p no:cacheprovider" 2>&1 | grep -v "Warning\|warnings.html\|Consider using\|return float\|^$" | tail -40 > gpurun_out/quick.log
cat gpurun_out/quick.log
if [ -n "$2" ]; then
  timeout 1200 python bench.py $2 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
  python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    print("value %.2f img/s  ms/step %.2f  e2e %.2f  launches %d  clocks %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
    print("roofline", {k: d['roofline'][k] for k in ('kernel','achieved','peak','frac')})
    print("matcher", d.get('matcher'))
    for k, v in d['kernel_breakdown'].items(): print("  %-22s %8.3f ms/step  %6.1f launches  share %.3f" % (k, v['ms_per_step'], v['launches_per_step'], v['share_of_step']))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_quick.err').read()[-3000:])
PY
fi
