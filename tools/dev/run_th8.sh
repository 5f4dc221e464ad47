#!/bin/bash
cd spe_b200
for u in 1 2 4; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DT8_UNROLL=$u -c csrc/talking_h8.cu -o _obj/talking_h8.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libspe_b200.so _obj/*.o -Xcompiler -fPIC -cudart static
  echo "== unroll $u"; (cd .. && TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1")
  echo "   bwd 10 warps:"; (cd .. && SPE_TH8_BWD_WARPS=10 TH8_CHECK_BIG=0 timeout 300 python tools/dev/th8_check.py 2>&1 | grep "N=1600.*s16=1")
done
