#!/bin/bash
# usage: tools/ab_kernels.sh <variant> [<variant> ...]   (variants = era_zkevm_test_harness_b200/libzkgpu_<variant>.so; "base" = libzkgpu.so)
# Per-kernel GPU time of ONE MainVM 2^20 proof for each build variant (ncu time-only pass over tools/profile_kernels.py prove).
set -u
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/era_zkevm_test_harness_b200/libzkgpu_$v.so; [ "$v" = base ] && lib=$PWD/era_zkevm_test_harness_b200/libzkgpu.so
  ZKGPU_LIB=$lib ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ab_$v.csv python tools/profile_kernels.py prove 20 > gpurun_out/ab_$v.log 2>&1
  echo "== $v: $(tail -1 gpurun_out/ab_$v.log)"; python tools/summarise_launches.py gpurun_out/ab_$v.csv | head -${AB_LINES:-9}
done
