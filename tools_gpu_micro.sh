#!/bin/bash
SPE_GEMM_NO_BN256=1 python tools/gemm_micro.py 2>&1 | grep dbg | sed 's/dbg=0/bn128/'
python tools/gemm_micro.py 2>&1 | grep dbg | sed 's/dbg=0/bn256/'
