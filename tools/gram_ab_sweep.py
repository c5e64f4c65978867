#!/usr/bin/env python
"""Teacher-forced A/B of the two Gram-path eigen-solvers on the small reference case (67 samples, T = 24, d = 10,
chi_max = 20): every bond of two sweeps is run from the SAME state with the Cholesky + register Jacobi solver and with
the shared-memory Jacobi on H; kept dimension, loss and the new two-site product are compared, then the sweep continues
from the default solver's result.  Separates a solver difference from the rounding-level divergence of free-running KLD
trajectories."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpstime_jl_b200 as m  # noqa: E402

rng = np.random.default_rng(7)
T, d, chi = 24, 10, 20
y = rng.integers(0, 2, 67)
t = np.arange(T)
X = np.where(y[:, None] == 0, np.sin(2 * np.pi * t / 24.0)[None, :], np.sin(2 * np.pi * t / 24.0 + 0.6)[None, :] * 0.8) + 0.25 * rng.standard_normal((67, T))
opts = m.MPSOptions(d=d, chi_max=chi, verbosity=-1)
Xs, _ = m.transform_train_data(X.T, opts)
Xs_sorted, _, ys, _, classes, counts = m.sort_by_class(Xs, X, y)
ctx = m.Context(0)
ctx.train_load_x(Xs_sorted, counts, d, chi)
cores = m.generate_starting_mps(4, T, d, 2, seed=1234)
topts = m.make_opts(chi_max=chi, eta=0.01)
order = ([(j, True) for j in range(T - 2, -1, -1)] + [(j, False) for j in range(T - 1)]) * 2
worst = {"product": 0.0, "loss_rel": 0.0, "chi_mismatch": 0, "gram_bonds": 0, "bonds": 0}
for j, gl in order:
    res = {}
    for flag in (1, 0):
        ctx.debug_set("SVD_GRAMREG", 1 - flag)
        ctx.set_cores(cores)
        ctx.build_env(True)
        ctx.build_env(False)
        lo, gn, k = ctx.bond_step(j, gl, topts)
        res[flag] = (lo, k, ctx.get_core(j), ctx.get_core(j + 1), ctx.debug_get("svd_path"))
    a, b = res[0], res[1]
    worst["bonds"] += 1
    if a[4] in (1, 2):
        worst["gram_bonds"] += 1
    if a[1] != b[1]:
        worst["chi_mismatch"] += 1
        print("chi mismatch at bond", j, gl, a[1], b[1], "path", a[4], flush=True)
    else:
        ein = "asmc,mtb->btasc" if gl else "asm,mtbc->btasc"
        worst["product"] = max(worst["product"], float(np.abs(np.einsum(ein, a[2], a[3]) - np.einsum(ein, b[2], b[3])).max()))
    worst["loss_rel"] = max(worst["loss_rel"], abs(a[0] - b[0]) / abs(b[0]))
    cores = ctx.get_cores()
ctx.debug_set("SVD_GRAMREG", 0)
print(json.dumps(worst))
json.dump(worst, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_gram_ab_sweep.json"), "w"))
