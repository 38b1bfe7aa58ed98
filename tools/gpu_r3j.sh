#!/bin/bash
# round-2, session 3j: last A/B (attempt counter, spline blocks per SM), then smoke + tests + bench on the final build (1 GPU)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product dg_nonstep product dg_nonstep; do run $v python tests/perf/time_dg.py 8192 3; done
} > gpurun_out/r3j_ab_dg.txt 2>&1
{
for v in product sp_mb4 sp_mb6; do run $v python tools/prof_spline.py 0.05 3; done
} > gpurun_out/r3j_ab_spline.txt 2>&1
unset B200CS_LIB
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3j_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3j_pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r3j_bench_n1.json 2> gpurun_out/r3j_bench_n1.err
timeout 600 python tests/perf/bench_configs.py > gpurun_out/r3j_configs_c1_c4.json 2> gpurun_out/r3j_configs.err
grep "lib=" gpurun_out/r3j_ab_dg.txt | cut -c1-120; cut -c1-120 gpurun_out/r3j_ab_spline.txt; cat gpurun_out/r3j_smoke.txt gpurun_out/r3j_pytest_gpu.txt | cut -c1-250; cut -c1-300 gpurun_out/r3j_bench_n1.json; grep -E '"ms"' gpurun_out/r3j_configs_c1_c4.json
