#!/bin/bash
# round-2, session 3m: fewer non-FP64 instructions in the double-gyre attempt (sign flip as IMAD, applied to r; controller literals from the constant bank)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product dg_flipadd dg_fe dg_fec product dg_fec; do run $v python tests/perf/time_dg.py 8192 3; done
} > gpurun_out/r3m_ab.txt 2>&1
grep -v "mismatch at" gpurun_out/r3m_ab.txt | cut -c1-170
