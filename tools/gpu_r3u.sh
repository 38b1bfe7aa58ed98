#!/bin/bash
# round-2 last session, multi-GPU check of the tiled kernels (one 8-GPU box): NCCL tests, bench at N = 8, 4, 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r3u_pytest_sharded.txt
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r3u_bench_n$n.json 2> gpurun_out/r3u_bench_n$n.err
done
cat gpurun_out/r3u_pytest_sharded.txt
for n in 8 4 2; do python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r3u_bench_n{n}.json"))
    print(n, "value %.1f M/s" % (d["value"] / 1e6), "e2e %.1f" % (d["e2e"]["value"] / 1e6), "ms %.2f" % d["ms_per_step"],
          "halo %.3f gather %.3f plan %.3f" % (d["halo_exchange_ms"], d["gather_ms"], d["planning_ms"]),
          "parity", d["parity"]["pass"], d["parity"]["step_mismatches"], "%.2e" % d["parity"]["max_rel_dx_matching"],
          "ridge %.2f ms %d pts" % (d["ridge_tail"]["ms"], d["ridge_tail"]["n_ridge_pts"]), "checksum", d["ftle_checksum"],
          "ftle_only", d["e2e"].get("ftle_only", {}).get("value"))
except Exception as e:
    print(n, "failed", e)
PY
done
tail -n 3 gpurun_out/r3u_bench_n8.err
