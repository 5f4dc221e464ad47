#!/bin/bash
python tools/dev/tf_fwd_check.py 2>&1 | tail -8
for d in ${SWEEP:-0 3 60 63}; do
  rm -f /tmp/prof.csv
  SPE_PROF_CSV=/tmp/prof.csv SPE_TF_DBG=$d TF_ONLY_BIG=1 TF_TIME=1 python tools/dev/tf_fwd_check.py 2>&1 | grep DBG
done
for n in ${NCH:-1 2 3 5}; do
  rm -f /tmp/prof.csv
  SPE_PROF_CSV=/tmp/prof.csv SPE_TF_NCHUNK=$n TF_ONLY_BIG=1 TF_TIME=1 python tools/dev/tf_fwd_check.py 2>&1 | grep DBG
done
