#!/bin/bash
# round-2, session 3d: A/B of the queue re-alignment interval / tile shapes, then the measurement set (1 GPU)
mkdir -p gpurun_out
V=$PWD/build/variants
run() { if [ "$1" = product ]; then unset B200CS_LIB; else export B200CS_LIB=$V/libb200cs_$1.so; fi; shift; timeout 300 "$@" 2>&1 | grep -v Warning; }
{
for v in product bk_qs3 bk_qs4 bk_t4 bk_t16; do
  run $v python tests/perf/time_bickley.py
  run $v python tools/prof_bickley.py 3 3
done
} > gpurun_out/r3d_ab_bickley.txt 2>&1
unset B200CS_LIB
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3d_pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r3d_bench_n1.json 2> gpurun_out/r3d_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r3d_dg_16384 -f python tools/run_dg.py 16384 2 > gpurun_out/r3d_ncu_dg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_queue_kernel -s 1 -c 1 \
    -o gpurun_out/r3d_bickley -f python tools/prof_bickley.py 1 2 > gpurun_out/r3d_ncu_bickley.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/r3d_spline -f python tools/prof_spline.py 0.05 2 > gpurun_out/r3d_ncu_spline.log 2>&1
timeout 600 python tests/perf/bench_configs.py > gpurun_out/r3d_configs_c1_c4.json 2> gpurun_out/r3d_configs.err
B200CS_LIB=$V/libb200cs_lavd_t1.so timeout 600 python tests/perf/bench_configs.py > gpurun_out/r3d_configs_c1_c4_lavd_t1.json 2>> gpurun_out/r3d_configs.err
python tools/prof_spline.py 0.05 3 > gpurun_out/r3d_time_spline.txt 2>&1
python tools/prof_bickley.py 1 3 > gpurun_out/r3d_time_bickley.txt 2>&1
cat gpurun_out/r3d_ab_bickley.txt gpurun_out/r3d_pytest_gpu.txt gpurun_out/r3d_time_spline.txt gpurun_out/r3d_time_bickley.txt | cut -c1-200; cut -c1-600 gpurun_out/r3d_bench_n1.json; tail -3 gpurun_out/r3d_bench_n1.err
