#!/bin/bash
# round-2, session 3h: fused LAVD kernel: paired output times and blocks per SM (1 GPU)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product lavd_p1mb3 lavd_p1mb4 lavd_p1mb5 lavd_p0 lavd_p0mb4 lavd_p0mb5; do run $v python tools/prof_lavd.py 3; done
} > gpurun_out/r3h_lavd.txt 2>&1
unset B200CS_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_atsize.py tests/test_gpu_config4.py -m gpu -q -k "lavd or config4 or c4" 2>&1 | tail -6 > gpurun_out/r3h_pytest_lavd.txt
cat gpurun_out/r3h_lavd.txt gpurun_out/r3h_pytest_lavd.txt | cut -c1-250
