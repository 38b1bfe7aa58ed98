#!/bin/bash
# round-2, session 3c: Bickley folded RHS, lockstep queue kernels, one-Newton controller (1 GPU)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product bk_nofold bk_ls640 bk_ls640s2 bk_ls640s1 bk_ls320 bk_ls512 bk_ls256 bk_ls640nf; do
  run $v python tests/perf/time_bickley.py
  run $v python tools/prof_bickley.py 3 3
done
} > gpurun_out/r3c_ab_bickley.txt 2>&1
{
for v in product bk_ls640 bk_ls320 bk_nofold bk_ls640nf; do run $v python tools/grid_hash.py; done
} > gpurun_out/r3c_hashes.txt 2>&1
{
for v in product dg_n1 product dg_n1; do run $v python tests/perf/time_dg.py 8192 3; done
} > gpurun_out/r3c_ab_dg.txt 2>&1
unset B200CS_LIB
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r3c_pytest_gpu.txt
cat gpurun_out/r3c_ab_bickley.txt gpurun_out/r3c_hashes.txt gpurun_out/r3c_ab_dg.txt gpurun_out/r3c_pytest_gpu.txt | cut -c1-200
