#!/bin/bash
# Builds A/B variants of the tcgen05 IIC adjoint (role split, poll back-off) into profiles/probes/variants/*.so.
# usage (CPU box): profiles/probes/iic_variants.sh name "-DCY_TC_NISS=2 ..." [name flags ...]
# on the GPU box:  for v in profiles/probes/variants/*.so; do cp $v contrast-you_b200/libcontrastyou_b200.so; python profiles/probes/iic_micro.py; done
set -e
here=$(cd "$(dirname "$0")" && pwd); csrc=$here/../../contrast-you_b200/csrc
mkdir -p $here/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
      -c $csrc/iic_bwd_tc.cu -o /tmp/iic_bwd_tc_$name.o
  objs=$(ls $csrc/build/*.o | grep -v iic_bwd_tc.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $here/variants/$name.so $objs /tmp/iic_bwd_tc_$name.o
  echo built $name
done
