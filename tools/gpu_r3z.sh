#!/bin/bash
# round-2, session 3z: fused LAVD kernel specialised on cubic-slab vorticity (A/B)
mkdir -p gpurun_out
V=$PWD/build/variants
{
python tools/prof_lavd.py 4
B200CS_LIB=$V/libb200cs_lavd_sp.so python tools/prof_lavd.py 4
python tools/prof_lavd.py 4
B200CS_LIB=$V/libb200cs_lavd_sp.so python tools/prof_lavd.py 4
} > gpurun_out/r3z_lavd.txt 2>&1
B200CS_LIB=$V/libb200cs_lavd_sp.so timeout 600 python -m pytest tests -m gpu -q -k "lavd or config4 or c4" 2>&1 | tail -4 >> gpurun_out/r3z_lavd.txt
cut -c1-130 gpurun_out/r3z_lavd.txt
