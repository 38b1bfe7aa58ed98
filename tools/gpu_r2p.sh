#!/bin/bash
# round-2 measurement session (1 GPU): bench line, ncu --set full of the integration kernel and of the TMA FTLE
# kernel at the bench size, launch list of the bench, configs 1-4, reference arm
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r2p_dg_16384 -f python tools/run_dg.py 16384 2 > gpurun_out/r2p_ncu_dg.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ftle_tma_kernel -c 1 \
    -o gpurun_out/r2p_ftle_tma_16384 -f python tools/time_ftle.py 16384 1 > gpurun_out/r2p_ncu_ftle.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2p_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2p_bench_under_ncu.log 2>&1
timeout 600 python tests/perf/bench_configs.py > gpurun_out/r2p_configs_c1_c4.json 2> gpurun_out/r2p_configs.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_reference.json 2>> gpurun_out/r2p_bench_n1.err
cat gpurun_out/r2p_bench_n1.json | cut -c1-1500; tail -3 gpurun_out/r2p_bench_n1.err gpurun_out/r2p_ncu_dg.log gpurun_out/r2p_configs.err; cat gpurun_out/r2p_bench_reference.json | cut -c1-900
