#!/bin/bash
# round-2, session 3o: bit-identity of the non-FP64 trims (hashes vs the previous arithmetic), tests, timings
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product prev dg_fe1; do run $v python tools/grid_hash.py; done
for v in product prev dg_fe1 product; do run $v python tests/perf/time_dg.py 8192 3; done
} > gpurun_out/r3o_ab.txt 2>&1
unset B200CS_LIB
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3o_pytest_gpu.txt
grep -v "mismatch at" gpurun_out/r3o_ab.txt | cut -c1-170; cat gpurun_out/r3o_pytest_gpu.txt
