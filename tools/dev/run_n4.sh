#!/bin/bash
mkdir -p gpurun_out
for ov in 1 0; do
  SPE_AR_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_n4_ov$ov.err | grep '^{' > gpurun_out/bench_n4_ov$ov.json
  echo "bench rc ${PIPESTATUS[0]}"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n4_ov$ov.json').read())
print('N=4 overlap=$ov', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
PY
done
