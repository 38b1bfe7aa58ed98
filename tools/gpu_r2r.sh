#!/bin/bash
mkdir -p gpurun_out
lscpu | grep -i "numa\|socket\|model name" > gpurun_out/r2r_lscpu.txt; nvidia-smi topo -m >> gpurun_out/r2r_lscpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29704 \
      bench.py --gpus 4 --steps 10 --warmup 3 --no-ridges > gpurun_out/r2r_bench_n4.json 2> gpurun_out/r2r_bench_n4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2r_bench_n4.json")); e = d["e2e"]
print("N=4 value %.1f e2e(with fm) %.1f (%.2f ms) ftle_only %.1f numa %s" % (d["value"] / 1e6, e["value"] / 1e6, e["ms_per_step"], e["ftle_only"]["value"] / 1e6, e.get("numa_node_of_rank0")))
PY
cat gpurun_out/r2r_lscpu.txt | head -30; tail -3 gpurun_out/r2r_bench_n4.err
