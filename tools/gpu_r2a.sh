#!/bin/bash
# round-2 first GPU session: sanity tests, FP64-conversion pipe micro-benchmark, ncu captures of the
# Bickley (config 2) and spline (config 3) integration kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
tools/ubench/xu_f64 > gpurun_out/r2a_xu_f64.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2a_pytest_gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r2a_bickley -f python tools/prof_bickley.py 1 2 > gpurun_out/r2a_ncu_bickley.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r2a_spline -f python tools/prof_spline.py 0.05 2 > gpurun_out/r2a_ncu_spline.log 2>&1
python tools/prof_bickley.py 1 3 > gpurun_out/r2a_time_bickley.txt 2>&1
python tools/prof_spline.py 0.05 3 > gpurun_out/r2a_time_spline.txt 2>&1
cat gpurun_out/r2a_xu_f64.txt gpurun_out/r2a_pytest_gpu.txt gpurun_out/r2a_time_bickley.txt gpurun_out/r2a_time_spline.txt
tail -n 3 gpurun_out/r2a_ncu_bickley.log gpurun_out/r2a_ncu_spline.log
