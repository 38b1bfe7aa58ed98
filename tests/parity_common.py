"""Shared pieces of the at-size parity tests (tests/test_gpu_parity_atsize.py), the calibration tool
(tests/perf/parity_floor.py) and bench.py's parity block: the BASELINE configurations' synthetic inputs
 and the flow-map comparison.

The comparison follows SURVEY section 8(d) "Parity gates": positions max|dx| <= 1e-8 x L over particles
whose (accepted, rejected) step counts equal the oracle's, the number of step-count mismatches
reported separately, FTLE relative L2 over pixels whose 5-point stencil touches no mismatch.
"""
import numpy as np


def compare_flowmaps(gpu, steps_gpu, ora, steps_o, L):
    """Per-particle |gpu - oracle| / L statistics; `steps_*` are [..., 2] (accepted, rejected)."""
    gpu, ora = np.asarray(gpu), np.asarray(ora)
    d = (np.abs(gpu - ora) / np.asarray(L)).max(axis=-1)
    same = (np.asarray(steps_gpu) == np.asarray(steps_o)).all(axis=-1)
    n = int(d.size)
    return {"particles": n, "step_mismatches": int((~same).sum()),
            "mismatch_fraction": float((~same).sum()) / max(n, 1),
            "max_rel_dx_matching": float(d[same].max()) if same.any() else 0.0,
            "max_rel_dx_all": float(d.max()) if n else 0.0,
            "over_1e-8_matching": int((d[same] > 1e-8).sum()),
            "p99_rel_dx": float(np.percentile(d, 99)) if n else 0.0,
            "median_rel_dx": float(np.median(d)) if n else 0.0}, same


def ftle_rel_l2(ft, fto, same=None):
    """Relative L2 error of an FTLE field; with `same` (per-particle step-count equality) the
    pixels whose 5-point stencil touches a step-count-mismatched particle are left out."""
    ft, fto = np.asarray(ft), np.asarray(fto)
    keep = np.ones(ft.shape, bool)
    if same is not None:
        bad = ~same
        touch = bad.copy()
        touch[1:] |= bad[:-1]
        touch[:-1] |= bad[1:]
        touch[:, 1:] |= bad[:, :-1]
        touch[:, :-1] |= bad[:, 1:]
        keep = ~touch
    den = np.linalg.norm(fto[keep])
    return float(np.linalg.norm((ft - fto)[keep]) / den) if den > 0 else 0.0


# ------------------------------------------------------------------ BASELINE configs 2-4: inputs

def bickley_grid(nx=2001, ny=601):
    """Config 2: Bickley jet on [0, 6.371 pi] x [-3, 3], T = +6 (plot_bickley_ftle.py:24)."""
    return np.linspace(0.0, 6.371 * np.pi, nx), np.linspace(-3.0, 3.0, ny)


def merra_axes():
    t = np.arange(720, dtype=np.float64)
    lon = -180.0 + 0.625 * np.arange(576, dtype=np.float64)
    lat = -90.0 + 0.5 * np.arange(361, dtype=np.float64)
    return t, lon, lat


def merra_field(xp, t, lon, lat, nt=None):
    """Config 3's synthetic MERRA-shaped velocity (km/h): eight Rossby-like modes, seed 0.
    `xp` is numpy or torch (the 720 x 576 x 361 field is built on the GPU when torch is passed)."""
    g = np.random.default_rng(0)
    if xp is np:
        Tm, LO, LA = np.meshgrid(t, np.deg2rad(lon), np.deg2rad(lat), indexing="ij")
        U, V = np.zeros_like(Tm), np.zeros_like(Tm)
    else:
        Tm, LO, LA = xp.meshgrid(t, xp.deg2rad(lon), xp.deg2rad(lat), indexing="ij")
        U, V = xp.zeros_like(Tm), xp.zeros_like(Tm)
    for _ in range(8):
        k, l = int(g.integers(1, 5)), int(g.integers(1, 4))
        ph, om = float(g.uniform(0, 6.28)), float(g.uniform(0.01, 0.05))
        au, av = float(g.uniform(5, 12)), float(g.uniform(3, 8))
        U += au * xp.cos(LA) * xp.sin(k * LO + om * Tm + ph) * xp.cos(l * LA)
        V += av * xp.cos(LA) * xp.cos(k * LO - om * Tm + ph) * xp.sin(2 * l * LA)
    return U, V


def merra_particles():
    """Config 3's FTLE grid: 0.2 degrees, lon [-100, 35] x lat [-5, 45] = 676 x 251."""
    return np.arange(-100, 35 + 0.1, 0.2), np.arange(-5, 45 + 0.1, 0.2)


def qge_field():
    """Config 4's QGE-shaped field: seeded stream function with psi = 0 on the walls of
    [0, 1] x [0, 2], 101 times; returns (t, x, y, U, V, vort)."""
    g = np.random.default_rng(0)
    for _ in range(8):      # keep the generator state of tests/perf/bench_configs.py (after config 3)
        g.integers(1, 5), g.integers(1, 4), g.uniform(0, 6.28), g.uniform(0.01, 0.05)
        g.uniform(5, 12), g.uniform(3, 8)
    xq, yq, tq = np.linspace(0, 1, 257), np.linspace(0, 2, 513), np.linspace(0, 1, 101)
    Tq, Xq, Yq = np.meshgrid(tq, xq, yq, indexing="ij")
    psi = np.zeros_like(Tq)
    for _ in range(6):
        k, l = int(g.integers(1, 4)), int(g.integers(1, 5))
        amp, om, ph = float(g.uniform(0.02, 0.06)), float(g.uniform(1, 6)), float(g.uniform(0, 6.28))
        psi += amp * np.sin(k * np.pi * Xq) * np.sin(l * np.pi * Yq / 2) * np.cos(om * Tq + ph)
    dxq, dyq = xq[1] - xq[0], yq[1] - yq[0]
    U = -np.gradient(psi, dyq, axis=2)
    V = np.gradient(psi, dxq, axis=1)
    vort = np.gradient(V, dxq, axis=1) - np.gradient(U, dyq, axis=2)
    return tq, xq, yq, U, V, vort


def c5_sample_rows(n=16384, blocks=6, rows_per_block=4, seed=5):
    """Config 5: a few contiguous row blocks of the n x n grid (both borders, the middle, random)."""
    rng = np.random.default_rng(seed)
    starts = [0, n // 2 - rows_per_block // 2, n - rows_per_block]
    starts += [int(v) for v in rng.integers(rows_per_block, n - 2 * rows_per_block, size=max(blocks - 3, 0))]
    return sorted(set(starts)), rows_per_block
