#!/bin/bash
mkdir -p gpurun_out
for v in product bk_rcp; do
  if [ "$v" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$PWD/build/variants/libb200cs_$v.so; fi
  timeout 200 python tests/perf/time_bickley.py 2>&1 | grep -v Warning
  timeout 200 python tools/prof_bickley.py 3 3 2>&1 | tail -1
done > gpurun_out/r2v_ab_bickley_rcp.txt 2>&1
cat gpurun_out/r2v_ab_bickley_rcp.txt
