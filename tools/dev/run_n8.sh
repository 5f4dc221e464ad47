#!/bin/bash
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_n8.err | grep '^{' > gpurun_out/bench_n8.json
echo "bench rc ${PIPESTATUS[0]}"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n8.json').read())
print('N=8', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
PY
tail -3 gpurun_out/bench_n8.err
