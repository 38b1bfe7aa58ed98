#!/bin/bash
# round-2, session 3i: final LAVD shape (slabs, four blocks per SM): tests, config timings, ncu (1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r3i_pytest_gpu.txt
python tools/prof_lavd.py 4 > gpurun_out/r3i_lavd.txt 2>&1
timeout 600 python tests/perf/bench_configs.py > gpurun_out/r3i_configs_c1_c4.json 2> gpurun_out/r3i_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lavd_flowmap_kernel -c 1 \
    -o gpurun_out/r3i_lavd -f python tools/prof_lavd.py 1 > gpurun_out/r3i_ncu_lavd.log 2>&1
cat gpurun_out/r3i_pytest_gpu.txt gpurun_out/r3i_lavd.txt | cut -c1-250; grep -E '"ms"|"C' gpurun_out/r3i_configs_c1_c4.json
