#!/bin/bash
timeout 300 python tools/dev/th16_check.py 2>&1 | grep "H=16"
for nt in 256 384; do echo "== bwd threads $nt"; SPE_TH16_BWD_THREADS=$nt timeout 300 python tools/dev/th16_check.py 2>&1 | grep "N=4150.*s16=1"; done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_extra_gpu.py tests/test_model_gpu.py -k "talking or h16" -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --config cfg4 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4', round(d['value'],2), round(d['ms_per_step'],1)); print({k:(round(v['ms_per_step'],2)) for k,v in d['kernel_breakdown'].items()})"
