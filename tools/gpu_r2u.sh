#!/bin/bash
mkdir -p gpurun_out
for v in product dg_fsal; do
    if [ "$v" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$PWD/build/variants/libb200cs_$v.so; fi
    timeout 300 python tests/perf/time_dg.py 8192 3 2>&1 | grep -v Warning
done > gpurun_out/r2u_ab_dg_nobranch.txt 2>&1
cat gpurun_out/r2u_ab_dg_nobranch.txt
