#!/bin/bash
# round-2, session 3y: block shapes / tile shapes re-checked with the final kernels
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product dg_64x10 dg_256x2 dg_t2 dg_t8 product; do run $v python tests/perf/time_dg.py 8192 3; done
for v in product sp_t2 sp_64 sp_256; do run $v python tools/prof_spline.py 0.05 3; done
} > gpurun_out/r3y_ab.txt 2>&1
grep "lib=\|grid" gpurun_out/r3y_ab.txt | cut -c1-130
