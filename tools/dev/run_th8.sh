#!/bin/bash
timeout 300 python tools/dev/th8_check.py 2>&1 | grep "H=8.*s16=1"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -k "talking or cfg2 or cfg1" -x -q -p no:cacheprovider 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step']); print({k:(round(v['ms_per_step'],2)) for k,v in d['kernel_breakdown'].items()})"
