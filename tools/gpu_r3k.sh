#!/bin/bash
# round-2, session 3k: the reference's CPU path on BASELINE configs 1-4 (host cores of the GPU box), LAVD re-check
mkdir -p gpurun_out
timeout 900 python tests/perf/cpu_configs.py > gpurun_out/r3k_cpu_configs.json 2> gpurun_out/r3k_cpu_configs.err
python tools/prof_lavd.py 3 > gpurun_out/r3k_lavd.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_atsize.py tests/test_gpu_config4.py tests/test_gpu_sharded.py -m gpu -q -k "lavd or config4 or c4" 2>&1 | tail -4 > gpurun_out/r3k_pytest_lavd.txt
cat gpurun_out/r3k_cpu_configs.err | tail -8; grep -E "points_per_s|cores" gpurun_out/r3k_cpu_configs.json; cat gpurun_out/r3k_lavd.txt gpurun_out/r3k_pytest_lavd.txt | cut -c1-200
