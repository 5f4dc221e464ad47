#!/bin/bash
for d in ${SWEEP:-0 63 127 191 255 319 511 64 67}; do
  rm -f /tmp/prof.csv
  SPE_PROF_CSV=/tmp/prof.csv SPE_TF_DBG=$d TF_ONLY_BIG=1 TF_TIME=1 python tools/dev/tf_fwd_check.py 2>&1 | grep DBG
done
