#!/bin/bash
# build variants of talking_fused.cu with different pipeline parameters and time the forward kernels (run on the GPU box)
cd /root/repo
for v in "$@"; do
  flags=$(echo $v | tr ',' ' ')
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $flags -c spe_b200/csrc/talking_fused.cu -o spe_b200/_obj/talking_fused.o 2>&1 | grep -E "error" 
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o spe_b200/libspe_b200.so spe_b200/_obj/*.o -Xcompiler -fPIC -cudart static
  rm -f /tmp/prof.csv
  echo "== $v"
  SPE_PROF_CSV=/tmp/prof.csv TF_ONLY_BIG=1 TF_TIME=1 python tools/dev/tf_fwd_check.py 2>&1 | grep -E "DBG|err" | tail -2
done
