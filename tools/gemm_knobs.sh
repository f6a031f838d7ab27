#!/bin/bash
# development aid: time GEMM shapes under the MUMPY_TC_* knobs
for what in gemm_fc1 gemm_fc2 gemm_qkv gemm_s0; do
  for cfg in "" "MUMPY_TC_DEBUG=1" "MUMPY_TC_DEBUG=2" "MUMPY_TC_BN=128"; do
    echo -n "$what [$cfg]: "
    env $cfg python tools/time_kernel.py $what
  done
done
