#!/bin/bash
# A/B of double-gyre kernel variants built by tools/build_variant.py, one process per variant.
#   tools/ab_dg.sh N variant [variant ...]   -> gpurun_out/ab_dg.txt
N=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
    B200CS_LIB=build/variants/libb200cs_$v.so timeout 300 python tests/perf/time_dg.py "$N" 3 2>&1 | grep -v Warning
done | tee gpurun_out/ab_dg.txt
