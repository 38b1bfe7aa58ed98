#!/bin/bash
# One GPU session: A/B timing + dense diagnostic, full GPU tests, ncu capture of the DG kernel,
# bench line, launch list, configs 1-4, time series.   usage: [NCU_N=16384] tools/gpu_round.sh TAG variant...
TAG=$1; shift
N=${NCU_N:-16384}
mkdir -p gpurun_out
tools/ab_dg2.sh 8192 product "$@"
cp gpurun_out/ab_dg2.txt gpurun_out/${TAG}_ab.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${TAG}_pytest_gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flowmap_kernel -s 1 -c 1 \
    -o gpurun_out/${TAG}_dg_$N -f python tools/run_dg.py $N 2 > gpurun_out/${TAG}_ncu.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 600 python tests/perf/bench_configs.py > gpurun_out/${TAG}_configs_c1_c4.json 2> gpurun_out/${TAG}_configs.err
timeout 300 python tools/time_series.py > gpurun_out/${TAG}_time_series.json 2> gpurun_out/${TAG}_time_series.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_ab.txt gpurun_out/${TAG}_pytest_gpu.txt gpurun_out/${TAG}_bench_n1.json gpurun_out/${TAG}_configs_c1_c4.json gpurun_out/${TAG}_time_series.json gpurun_out/${TAG}_bench_reference.json
tail -n 3 gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_ncu.log gpurun_out/${TAG}_configs.err
