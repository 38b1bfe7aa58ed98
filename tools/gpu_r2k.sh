#!/bin/bash
mkdir -p gpurun_out
{
B200CS_FTLE_NO_TMA=1 python tools/time_ftle.py 16384 10
python tools/time_ftle.py 16384 10
for v in ft_tr8s4 ft_tr8s3 ft_tr4s3 ft_tr4s6 ft_tr2s4 ft_tr2s8; do
  B200CS_LIB=$PWD/build/variants/libb200cs_$v.so python tools/time_ftle.py 16384 10
done
} > gpurun_out/r2k_ftle_shapes.txt 2>&1
cat gpurun_out/r2k_ftle_shapes.txt
