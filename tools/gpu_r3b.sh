#!/bin/bash
# round-2, session 3b: queue kernels (lane-level work fetch) A/B + new tests (1 GPU)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product old_shape bk_noq bk_q4 bk_q_t4 bk_q_t1 sp_q dg_q; do run $v python tools/grid_hash.py; done
} > gpurun_out/r3b_hashes.txt 2>&1
{
for v in product old_shape bk_noq bk_q4 bk_q_t4 bk_q_t1; do
  run $v python tests/perf/time_bickley.py
  run $v python tools/prof_bickley.py 3 3
done
} > gpurun_out/r3b_ab_bickley.txt 2>&1
{
for v in product dg_q old_shape; do run $v python tests/perf/time_dg.py 8192 3; done
} > gpurun_out/r3b_ab_dg.txt 2>&1
{
for v in product sp_q old_shape; do run $v python tools/prof_spline.py 0.05 3; done
} > gpurun_out/r3b_ab_spline.txt 2>&1
unset B200CS_LIB
timeout 600 python -m pytest tests/test_gpu_queue.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r3b_pytest_queue.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_queue_kernel -c 1 \
    -o gpurun_out/r3b_bickley_queue -f python tools/prof_bickley.py 1 2 > gpurun_out/r3b_ncu_bickley.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r3b_pytest_gpu.txt
cat gpurun_out/r3b_hashes.txt gpurun_out/r3b_ab_bickley.txt gpurun_out/r3b_ab_dg.txt gpurun_out/r3b_pytest_queue.txt gpurun_out/r3b_pytest_gpu.txt | cut -c1-200; cut -c1-110 gpurun_out/r3b_ab_spline.txt
