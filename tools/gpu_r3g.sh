#!/bin/bash
# round-2, session 3g: LAVD with time-collapsed vorticity slabs (A/B against the 64-tap evaluator), tests (1 GPU)
mkdir -p gpurun_out
{
python tools/prof_lavd.py 3
B200CS_LAVD_NO_SLABS=1 python tools/prof_lavd.py 3
} > gpurun_out/r3g_lavd.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r3g_lavd_launches.csv python tools/prof_lavd.py 1 > /dev/null 2>&1
B200CS_LAVD_NO_SLABS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r3g_lavd_launches_noslabs.csv python tools/prof_lavd.py 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lavd_flowmap_kernel -c 1 \
    -o gpurun_out/r3g_lavd -f python tools/prof_lavd.py 1 > gpurun_out/r3g_ncu_lavd.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r3g_pytest_gpu.txt
cat gpurun_out/r3g_lavd.txt gpurun_out/r3g_pytest_gpu.txt | cut -c1-300
python tools/summarize_launches.py gpurun_out/r3g_lavd_launches.csv | head -12
python tools/summarize_launches.py gpurun_out/r3g_lavd_launches_noslabs.csv | head -12
