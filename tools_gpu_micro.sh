#!/bin/bash
for d in 0 1 2 3 4 7 8 15; do SPE_GEMM_DBG=$d python tools/gemm_micro.py 2>&1 | grep dbg; done
