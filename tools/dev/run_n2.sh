#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_nccl_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
for ov in 1 0; do
  SPE_AR_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench_n2_ov$ov.err | grep '^{' > gpurun_out/bench_n2_ov$ov.json
  echo "bench rc ${PIPESTATUS[0]}"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_ov$ov.json').read())
print('overlap=$ov', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29813 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --no-ref-gpu 2>/dev/null | cut -c 1-200
