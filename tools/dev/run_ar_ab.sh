#!/bin/bash
# 2-GPU: NCCL parity test, then bench with and without the overlapped bucketed all-reduce
mkdir -p gpurun_out
python -m pytest tests/test_dist_nccl_gpu.py -x -q 2>&1 | tail -5
for ov in 1 0; do
  SPE_AR_OVERLAP=$ov python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | grep '^{' > gpurun_out/bench_n2_ov$ov.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_ov$ov.json').read())
print('overlap=$ov', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
PY
done
python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' > gpurun_out/bench_n1.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read()); print('n1', d['value'], d['ms_per_step'], d['e2e']['value'])"
