#!/bin/bash
# round 2, pass m: CSR build at 256 / 512 / 1024 threads per CTA with the two-deep pipeline; MPRGP n=4096 ncu stalls
out=gpurun_out; tag=r02m
mkdir -p $out
for t in 256 512 1024; do
  lib=$PWD/ccqppy_b200/csrc/variants/libccqp_csr$t.so; [ $t = 1024 ] && lib=$PWD/ccqppy_b200/csrc/libccqp_b200.so
  for g in 1 2 4 8; do echo "threads=$t CSR_GROUP=$g"; CCQP_B200_LIB=$lib CCQP_CSR_GROUP=$g timeout 300 python tools/profile_csr.py 2>&1 | grep "csr gemv\|max rel"; done
  echo "threads=$t solve"; CCQP_B200_LIB=$lib CCQP_DEBUG_TIMING=1 timeout 300 python tools/profile_csr.py --solve 2>&1 | grep -v Warn | tail -2
done
