#!/bin/bash
# round-2, session 3e: double-gyre slopes in shared memory (A/B), launch list of the bench, tests on the final build (1 GPU)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product dg_ks6 dg_ks7 product; do run $v python tests/perf/time_dg.py 8192 3; done
} > gpurun_out/r3e_ab_dg.txt 2>&1
unset B200CS_LIB
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3e_pytest_gpu.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r3e_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r3e_bench_under_ncu.log 2>&1
timeout 600 python tests/perf/bench_configs.py > gpurun_out/r3e_configs_c1_c4.json 2> gpurun_out/r3e_configs.err
python tools/prof_bickley.py 1 3 > gpurun_out/r3e_time_bickley.txt 2>&1
python tools/prof_bickley.py 3 3 >> gpurun_out/r3e_time_bickley.txt 2>&1
timeout 300 python tools/time_series.py > gpurun_out/r3e_time_series.json 2> gpurun_out/r3e_time_series.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3e_bench_reference.json 2> gpurun_out/r3e_bench_reference.err
grep -v "mismatch at\|particles with" gpurun_out/r3e_ab_dg.txt | cut -c1-170; cat gpurun_out/r3e_pytest_gpu.txt gpurun_out/r3e_time_bickley.txt; cut -c1-300 gpurun_out/r3e_bench_reference.json
