#!/bin/bash
# round validation on one B200: the GPU test suite, smoke(), default bench, reference arm (CPU + informational GPU eager), cfg4 bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.html\|Consider using\|return float\|^$" | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python bench.py --config cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
echo "--- tests"; cat gpurun_out/pytest_gpu.log
echo "--- smoke"; tail -3 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench_default", "bench_cfg4"):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, "value %.2f img/s  ms/step %.2f  e2e %.2f  launches %d  clocks %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
        print("  roofline", {k: d['roofline'][k] for k in ('kernel','achieved','peak','frac')}, d['roofline']['attention_gemm']['frac'])
        print("  cpu", d['cpu_baseline'])
    except Exception as e:
        print(f, "parse failed", e); print(open('gpurun_out/%s.err' % f).read()[-2000:])
try:
    d = json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
    print("reference", d['value'], d['cpu_baseline'], d.get('reference_gpu_eager'))
except Exception as e:
    print("reference parse failed", e); print(open('gpurun_out/bench_reference.err').read()[-2000:])
PY
