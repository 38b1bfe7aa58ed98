#!/bin/bash
# round-2, session 3s: dense-output theta with one reciprocal per step (A/B on config 4 and the dense tests)
mkdir -p gpurun_out
V=$PWD/build/variants
{
python tools/prof_lavd.py 4
B200CS_LIB=$V/libb200cs_lavd_rcp.so python tools/prof_lavd.py 4
python tools/prof_lavd.py 4
B200CS_LIB=$V/libb200cs_lavd_rcp.so python tools/prof_lavd.py 4
} > gpurun_out/r3s_lavd.txt 2>&1
B200CS_LIB=$V/libb200cs_lavd_rcp.so timeout 600 python -m pytest tests -m gpu -q -k "lavd or dense or flowmap_n or golden or config4 or c4" 2>&1 | tail -5 > gpurun_out/r3s_pytest.txt
cut -c1-120 gpurun_out/r3s_lavd.txt; cat gpurun_out/r3s_pytest.txt
