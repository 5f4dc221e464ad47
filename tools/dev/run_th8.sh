#!/bin/bash
for cfgs in "3 8" "3 10" "2 10"; do set -- $cfgs; echo "== fwd ctas $1 warps $2"; SPE_TH8_FWD_CTAS=$1 SPE_TH8_FWD_WARPS=$2 TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1"; done
