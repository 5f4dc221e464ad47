#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/gemm_shapes.csv
SPE_PROF_CSV=gpurun_out/gemm_shapes.csv timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.err
python - <<'PY'
import csv, collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for fam, tag, work, ms in csv.reader(open('gpurun_out/gemm_shapes.csv')):
    if fam != '0': continue
    a = agg[tag]; a[0] += 1; a[1] += float(ms); a[2] += float(work)
tot = sum(a[1] for a in agg.values())
print("total gemm ms (3 steps)", tot)
for tag, (n, ms, w) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-44s n=%4d ms=%8.3f (%.1f%%) avg=%.3f ms  %.1f TF/s" % (tag, n, ms, 100*ms/tot, ms/n, w/ms/1e9 if ms else 0))
PY
