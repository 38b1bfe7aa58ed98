#!/bin/bash
# round-2, session 3v: Bickley sincos as fast-path-then-fix-up (A/B)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product bk_fix; do run $v python tools/grid_hash.py; done
for v in product bk_fix product bk_fix; do run $v python tools/prof_bickley.py 1 3; run $v python tools/prof_bickley.py 3 3; done
run bk_fix python -m pytest tests/test_gpu_parity.py tests/test_gpu_queue.py -m gpu -q -k "bickley or rhs or queue or launch" 
} > gpurun_out/r3v_ab.txt 2>&1
cut -c1-170 gpurun_out/r3v_ab.txt | tail -16
