#!/bin/bash
# usage: tools_gpu_run.sh <logname> <pytest -k expr or file> ...   runs each arg as a separate pytest process
mkdir -p gpurun_out
log=gpurun_out/$1.log; shift
: > $log
for k in "$@"; do
  echo "=== $k" >> $log
  eval "timeout 900 python -m pytest $k -q -x -p no:cacheprovider" 2>&1 | grep -v "Warning\|warnings.html\|Consider using\|return float\|^$" | tail -40 >> $log
done
tail -200 $log
