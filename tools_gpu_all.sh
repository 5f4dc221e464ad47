#!/bin/bash
# full GPU validation: kernel tests, model tests, smoke, short bench with per-shape GEMM profile
mkdir -p gpurun_out
./tools_gpu_run.sh all "tests/test_kernels_gpu.py -k gemm" "tests/test_kernels_gpu.py -k 'not gemm'" "tests/test_model_gpu.py -k tiny" "tests/test_model_gpu.py -k 'not tiny'" > /dev/null 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
rm -f gpurun_out/gemm_shapes.csv
SPE_PROF_CSV=gpurun_out/gemm_shapes.csv timeout 1200 python bench.py --steps 4 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "--- tests"; grep -E "^===|passed|failed|Error|assert " gpurun_out/all.log | cut -c1-250 | head -60
echo "--- smoke"; tail -2 gpurun_out/smoke.log
echo "--- bench"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print("value %.2f img/s  ms/step %.2f  e2e %.2f  launches %d  clocks %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
    print("roofline", {k: d['roofline'][k] for k in ('kernel','achieved','peak','frac')})
    for k, v in d['kernel_breakdown'].items(): print("  %-22s %8.3f ms/step  %6.1f launches  share %.3f" % (k, v['ms_per_step'], v['launches_per_step'], v['share_of_step']))
    print("hbm GB/s", d['roofline']['hbm_families_gbs'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
python - <<'PY'
import csv, collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
try:
    for fam, tag, work, ms in csv.reader(open('gpurun_out/gemm_shapes.csv')):
        if fam != '0': continue
        a = agg[tag]; a[0] += 1; a[1] += float(ms); a[2] += float(work)
    tot = sum(a[1] for a in agg.values())
    print("total gemm ms (profiled steps)", tot)
    for tag, (n, ms, w) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print("%-44s n=%4d ms=%8.3f (%.1f%%) avg=%.3f ms  %.1f TF/s" % (tag, n, ms, 100*ms/tot, ms/n, w/ms/1e9 if ms else 0))
except Exception as e: print(e)
PY
