#!/bin/bash
python tools/gemm_micro.py 2>&1 | grep dbg | sed 's/^/bn128 /'
SPE_GEMM_BN256=1 python tools/gemm_micro.py 2>&1 | grep dbg | sed 's/^/bn256 /'
SPE_GEMM_BN256=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bn256 bench', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernel_breakdown'].items()})"
