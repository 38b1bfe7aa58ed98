#!/bin/bash
# round-2, session 3r: step-size clamps as compare + select / branch-free rsqrt of the error norm (A/B)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product dg_tc dg_tr dg_trc; do run $v python tools/grid_hash.py; done
for v in product dg_tc dg_tr dg_trc product dg_tc; do run $v python tests/perf/time_dg.py 8192 3; done
} > gpurun_out/r3r_ab.txt 2>&1
grep -v "mismatch at\|particles with\|C1 parity" gpurun_out/r3r_ab.txt | cut -c1-170
