#!/bin/bash
mkdir -p gpurun_out
for ov in 1; do
  echo "=== graph=1 overlap=$ov"
  SPE_TEST_DUMP_AFTER=60 SPE_TEST_GRAPH=1 SPE_AR_OVERLAP=$ov NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29840+ov)) tests/_nccl_parity_worker.py > gpurun_out/ar_dbg_$ov.log 2>&1
  echo "rc $?"
  grep "rank [01]\|Timeout\|File" gpurun_out/ar_dbg_$ov.log | tail -20
done
for ov in 1 0; do
  SPE_AR_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_n2_ov$ov.err | grep '^{' > gpurun_out/bench_n2_ov$ov.json
  echo "bench rc ${PIPESTATUS[0]}"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_ov$ov.json').read())
print('overlap=$ov', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
PY
done
