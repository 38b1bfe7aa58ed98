#!/bin/bash
# round-2 final measurement session (1 GPU)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r2z_dg_16384 -f python tools/run_dg.py 16384 2 > gpurun_out/r2z_ncu_dg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r2z_bickley -f python tools/prof_bickley.py 1 2 > gpurun_out/r2z_ncu_bickley.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r2z_spline -f python tools/prof_spline.py 0.05 2 > gpurun_out/r2z_ncu_spline.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2z_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2z_bench_under_ncu.log 2>&1
timeout 600 python tests/perf/bench_configs.py > gpurun_out/r2z_configs_c1_c4.json 2> gpurun_out/r2z_configs.err
python tools/prof_spline.py 0.05 3 > gpurun_out/r2z_time_spline.txt 2>&1
python tools/prof_bickley.py 1 3 > gpurun_out/r2z_time_bickley.txt 2>&1
timeout 300 python tools/time_series.py > gpurun_out/r2z_time_series.json 2> gpurun_out/r2z_time_series.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2z_pytest_gpu.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2>> gpurun_out/r2z_bench_n1.err
cut -c1-400 gpurun_out/r2z_bench_n1.json; cat gpurun_out/r2z_pytest_gpu.txt gpurun_out/r2z_time_spline.txt gpurun_out/r2z_time_bickley.txt | cut -c1-200
