#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor_ridges.py tests/test_gpu_sharded.py -q -x 2>&1 | tail -4 > gpurun_out/r2m_pytest.txt
python tools/bench_tensor.py 16384 > gpurun_out/r2m_bench_tensor_16384.json 2> gpurun_out/r2m_bench_tensor.err
python - <<'PY' > gpurun_out/r2m_fused.txt 2>&1
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from numbacs_b200.diagnostics import C_eig_2D, ftle_from_eig
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import flowmap_grid_2D
from numbacs_b200.extraction import ftle_ridge_pts
n = 16384
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
x = torch.linspace(0, 2, n, dtype=torch.float64, device="cuda"); y = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
fm = flowmap_grid_2D(f, 0., -10., x, y, p, device_out=True)
dx, dy = 2.0 / (n - 1), 1.0 / (n - 1)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best, out
t1, (vals, vecs, ft) = timed(lambda: C_eig_2D(fm, dx, dy, ftle_T=-10.0))
t2, (v2, e2) = timed(lambda: C_eig_2D(fm, dx, dy))
t3, ft2 = timed(lambda: ftle_from_eig(v2[:, :, 1], -10.0))
print(f"C_eig_2D fused with FTLE {t1:.3f} ms; separate: C_eig_2D {t2:.3f} + ftle_from_eig {t3:.3f} ms; identical {bool(torch.equal(ft, ft2))}")
t4, rp = timed(lambda: ftle_ridge_pts(ft, vecs[:, :, :, 1], x, y, sdd_thresh=10.0, percentile=0))
print(f"ftle_ridge_pts {t4:.3f} ms, {rp.shape[0]} points; tail total {t1 + t4:.3f} ms")
PY
cat gpurun_out/r2m_pytest.txt gpurun_out/r2m_fused.txt; tail -3 gpurun_out/r2m_bench_tensor.err
