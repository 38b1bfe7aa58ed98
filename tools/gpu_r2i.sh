#!/bin/bash
# round-2: ordered-ridge GPU tests + double-gyre kernel A/B (RHS trims, fast controller, blocks per SM)
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_gpu_tensor_ridges.py tests/test_ridge_link_cpu.py -q -x -s 2>&1 | tail -8 > gpurun_out/r2i_pytest_ridges.txt
for v in dg_t0c0 dg_t1c0 dg_t0c1 product dg_mb6 dg_mb7 dg_mb8; do
    if [ "$v" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$PWD/build/variants/libb200cs_$v.so; fi
    timeout 300 python tests/perf/time_dg.py 8192 3 2>&1 | grep -v Warning
done > gpurun_out/r2i_ab_dg.txt 2>&1
cat gpurun_out/r2i_pytest_ridges.txt gpurun_out/r2i_ab_dg.txt
