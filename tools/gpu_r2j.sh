#!/bin/bash
# round-2: TMA-fed FTLE kernel -- correctness against the register-rolling kernel and the oracle, timing
mkdir -p gpurun_out
{
python tools/time_ftle.py 16384 3
true
true
true
python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import oracle as O
from numbacs_b200.diagnostics import ftle_grid_2D, ftle_slab_2D
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import flowmap_grid_2D
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
for (nx, ny) in ((700, 1000), (513, 257), (2049, 1031), (64, 252), (1000, 253)):
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    fm = flowmap_grid_2D(f, 0., -10., x, y, p)
    dx, dy = x[1] - x[0], y[1] - y[0]
    mask = np.random.default_rng(1).random((nx, ny)) < 0.05
    for m in (None, mask):
        a = ftle_grid_2D(fm, -10., dx, dy, mask=m)
        ref = O.ftle_grid_2D(fm, -10., dx, dy, mask=m)
        err = np.abs(a - ref).max()
        zeros_ok = np.array_equal(a == 0, ref == 0)
        print(nx, ny, "mask" if m is not None else "nomask", "max abs diff vs oracle %.2e" % err, "zero pattern equal", zeros_ok,
              "borders", float(np.abs(a[0]).max() + np.abs(a[-1]).max() + np.abs(a[:, 0]).max() + np.abs(a[:, -1]).max()))
    # slab semantics: rows 100..400 of the grid with one halo row each side == the same rows of the full field
    if nx > 500:
        full = ftle_grid_2D(fm, -10., dx, dy)
        sl = ftle_slab_2D(fm[99:402], -10., dx, dy, (1, 1))
        print("   slab rows identical:", np.array_equal(sl, full[100:401]))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_atsize.py tests/test_gpu_tensor_ridges.py -q -k "ftle or c5 or c1 or c_eig" 2>&1 | tail -4
} > gpurun_out/r2j_ftle_tma.txt 2>&1
cat gpurun_out/r2j_ftle_tma.txt
