#!/bin/bash
mkdir -p gpurun_out
echo "== talking_h8.cu (default)"; timeout 300 python tools/dev/th8_check.py 2>&1 | grep "H=8.*s16=1"
echo "== rowwise.cu"; SPE_TH8_STREAM=0 TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1"
for nw in 8 16; do echo "== warps $nw"; SPE_TH8_FWD_WARPS=$nw SPE_TH8_BWD_WARPS=$nw TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1"; done
timeout 600 python -m pytest tests/test_kernels_gpu.py -k talking -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step']); print({k:(round(v['ms_per_step'],2)) for k,v in d['kernel_breakdown'].items()})"
