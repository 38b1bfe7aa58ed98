#!/bin/bash
# A/B of Bickley-jet kernel variants (tools/build_tu_variant.sh NAME flowmap_bickley ...): config-2 timing with
# its parity figures, then the Bickley / RHS parity tests of the GPU suite against the same library.
#   tools/ab_bickley.sh variant...   ("product" = in-tree library)  -> gpurun_out/ab_bickley.txt
mkdir -p gpurun_out
for v in "$@"; do
    if [ "$v" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$PWD/build/variants/libb200cs_$v.so; fi
    timeout 200 python tests/perf/time_bickley.py 2>&1 | grep -v Warning
    timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "bickley or rhs_ or velocity" 2>&1 | tail -4
done > gpurun_out/ab_bickley.txt 2>&1
cat gpurun_out/ab_bickley.txt
