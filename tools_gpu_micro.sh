#!/bin/bash
SPE_GEMM_GENERIC_EPILOGUE=1 python tools/gemm_micro.py 2>&1 | grep dbg | sed 's/^/generic /'
python tools/gemm_micro.py 2>&1 | grep dbg | sed 's/^/fast    /'
