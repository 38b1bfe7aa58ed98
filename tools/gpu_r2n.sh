#!/bin/bash
# round-2 spline / Bickley A/B: stage slopes in shared memory, out-of-line RHS, free-running blocks
mkdir -p gpurun_out
for v in product sp_k5 sp_k6 sp_k7 sp_k6i sp_k4i sp_kls; do
  if [ "$v" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$PWD/build/variants/libb200cs_$v.so; fi
  echo "== $v"; timeout 300 python tools/prof_spline.py 0.05 3 2>&1 | tail -1 | cut -c1-120
done > gpurun_out/r2n_ab_spline.txt 2>&1
for v in product bk_k6 bk_k7; do
  if [ "$v" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$PWD/build/variants/libb200cs_$v.so; fi
  timeout 200 python tests/perf/time_bickley.py 2>&1 | grep -v Warning
done >> gpurun_out/r2n_ab_spline.txt 2>&1
cat gpurun_out/r2n_ab_spline.txt
