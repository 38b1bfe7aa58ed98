#!/bin/bash
# round-2, session 3t: validation of the final defaults: smoke, tests, bench, ncu of the integration kernel, configs
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3t_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3t_pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r3t_bench_n1.json 2> gpurun_out/r3t_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r3t_dg_16384 -f python tools/run_dg.py 16384 2 > gpurun_out/r3t_ncu_dg.log 2>&1
timeout 600 python tests/perf/bench_configs.py > gpurun_out/r3t_configs_c1_c4.json 2> gpurun_out/r3t_configs.err
python tools/prof_bickley.py 1 3 > gpurun_out/r3t_time.txt 2>&1; python tools/prof_bickley.py 3 3 >> gpurun_out/r3t_time.txt 2>&1; python tools/prof_spline.py 0.05 3 >> gpurun_out/r3t_time.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r3t_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r3t_bench_under_ncu.log 2>&1
cat gpurun_out/r3t_smoke.txt gpurun_out/r3t_pytest_gpu.txt | cut -c1-250; cut -c1-130 gpurun_out/r3t_time.txt; cut -c1-400 gpurun_out/r3t_bench_n1.json; grep -E '"ms"' gpurun_out/r3t_configs_c1_c4.json
