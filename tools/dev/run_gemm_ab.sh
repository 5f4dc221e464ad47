#!/bin/bash
for v in "" "SPE_GEMM_DIRECT_EPILOGUE=1" "SPE_GEMM_BN256=1"; do
  echo "== $v"
  env $v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value'],1), round(d['ms_per_step'],2)); print({k:(round(v['ms_per_step'],2)) for k,v in d['kernel_breakdown'].items()})"
done
