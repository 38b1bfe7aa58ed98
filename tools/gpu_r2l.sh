#!/bin/bash
mkdir -p gpurun_out
{
echo "# register-rolling kernel, lean sqrt + no I2F (product, TMA off):"; B200CS_FTLE_NO_TMA=1 python tools/time_ftle.py 16384 10
echo "# same with the IEEE sqrt:"; B200CS_FTLE_NO_TMA=1 B200CS_LIB=$PWD/build/variants/libb200cs_ft_ieee.so python tools/time_ftle.py 16384 10
echo "# TMA kernel (product):"; python tools/time_ftle.py 16384 10
} > gpurun_out/r2l_ftle.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ftle_tma_kernel -c 1 -o gpurun_out/r2l_ftle_tma -f python tools/time_ftle.py 16384 1 > gpurun_out/r2l_ncu1.log 2>&1
B200CS_FTLE_NO_TMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ftle_kernel -c 1 -o gpurun_out/r2l_ftle_roll -f python tools/time_ftle.py 16384 1 > gpurun_out/r2l_ncu2.log 2>&1
cat gpurun_out/r2l_ftle.txt; tail -2 gpurun_out/r2l_ncu1.log gpurun_out/r2l_ncu2.log
