#!/bin/bash
echo "== fwd 5 warps x 4 ctas, bwd 5 warps"; SPE_TH8_FWD_CTAS=4 SPE_TH8_FWD_WARPS=5 SPE_TH8_BWD_WARPS=5 TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1"
echo "== fwd 5 warps x 3 ctas"; SPE_TH8_FWD_CTAS=3 SPE_TH8_FWD_WARPS=5 TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1"
echo "== fwd 8 warps x 3 ctas nbuf3, bwd 8"; SPE_TH8_FWD_NBUF=3 TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step']); print({k:(round(v['ms_per_step'],2)) for k,v in d['kernel_breakdown'].items()})"
