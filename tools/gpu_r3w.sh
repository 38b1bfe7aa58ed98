#!/bin/bash
# round-2, session 3w: last check of the committed tree (clean build): smoke, tests, short bench
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3w_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r3w_pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r3w_bench_n1.json 2> gpurun_out/r3w_bench_n1.err
python tools/grid_hash.py > gpurun_out/r3w_hash.txt 2>&1
cat gpurun_out/r3w_smoke.txt gpurun_out/r3w_pytest_gpu.txt gpurun_out/r3w_hash.txt | cut -c1-220; cut -c1-200 gpurun_out/r3w_bench_n1.json
