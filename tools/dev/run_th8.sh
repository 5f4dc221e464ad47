#!/bin/bash
for c in 3 4; do for nb in 2 3; do echo "== fwd ctas $c nbuf $nb"; SPE_TH8_FWD_CTAS=$c SPE_TH8_FWD_NBUF=$nb TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1"; done; done
echo "== ctas 3 nw 8"; SPE_TH8_FWD_CTAS=3 SPE_TH8_FWD_NBUF=2 SPE_TH8_FWD_WARPS=8 TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1"
