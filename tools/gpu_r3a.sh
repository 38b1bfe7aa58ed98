#!/bin/bash
# round-2, session 3a: A/B of lane tiling, rolled Bickley stages, aux scheduling, fast hinit (1 GPU)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product dg_tile4 dg_tile8 bk_tile4 bk_tile8 bk_roll bk_roll_t4; do run $v python tools/grid_hash.py; done
} > gpurun_out/r3a_hashes.txt 2>&1
{
for v in product dg_tile4 dg_tile8 dg_aux1 dg_aux2 dg_aux3 dg_hfast product; do run $v python tests/perf/time_dg.py 8192 3; done
} > gpurun_out/r3a_ab_dg.txt 2>&1
{
for v in product bk_tile4 bk_tile8 bk_roll bk_roll4 bk_roll_t4 bk_roll4_t4; do
  run $v python tests/perf/time_bickley.py
  run $v python tools/prof_bickley.py 3 3
done
} > gpurun_out/r3a_ab_bickley.txt 2>&1
{
for v in product sp_tile4 sp_tile8; do run $v python tools/prof_spline.py 0.05 3; done
} > gpurun_out/r3a_ab_spline.txt 2>&1
cat gpurun_out/r3a_hashes.txt gpurun_out/r3a_ab_dg.txt gpurun_out/r3a_ab_bickley.txt gpurun_out/r3a_ab_spline.txt | cut -c1-220
