#!/bin/bash
# usage: tools/ab_dg2.sh N variant...  ("product" = the in-tree library); timing + dense-output diagnostic per variant
N=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
    if [ "$v" = product ]; then unset B200CS_LIB; else export B200CS_LIB=build/variants/libb200cs_$v.so; fi
    timeout 300 python tests/perf/time_dg.py "$N" 3 2>&1 | grep -v Warning
    timeout 300 python tests/perf/diag_dense.py 2>&1 | grep -v Warning
done > gpurun_out/ab_dg2.txt 2>&1
