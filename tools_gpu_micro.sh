#!/bin/bash
python tools/gemm_micro.py 2>&1 | grep dbg
