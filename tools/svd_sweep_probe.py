"""K5 at a given bond shape inside real sweeps (the SVD cost does not depend on N): ms per split, mean subspace
iterations, second rounds and Jacobi fallbacks for a list of switch settings.
    python tools/svd_sweep_probe.py [d chi T N nsweeps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpstime_jl_b200 as m
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import make_data, WORKLOADS

d, chi, T, N, ns = (int(a) for a in (sys.argv[1:6] + [16, 64, 24, 8192, 3][len(sys.argv) - 1:]))
w = dict(WORKLOADS["C"])
X, y = make_data(N, T, 5, w)
opts = m.MPSOptions(d=d, chi_max=chi)
Xs, _ = m.transform_train_data(X.T, opts)
Xs, _, ys, _, classes, counts = m.sort_by_class(Xs, X, y)
cores0 = m.generate_starting_mps(4, T, d, 2, seed=1234)
topts = m.make_opts(chi_max=chi, eta=0.01)
ctx = m.Context(0)
SETTINGS = [{}] if os.environ.get("ONLY_DEFAULT") else [{}, {"SVD_IT": 6}, {"SVD_IT": 5}, {"SVD_IT": 4}, {"SVD_OVS": -16}, {"SVD_OVS": -32}, {"SVD_NOHALF": 1}]
for st in SETTINGS:
    for k in ("SVD_IT", "SVD_OVS", "SVD_NOHALF"):
        ctx.debug_set(k, st.get(k, 0))
    ctx.train_load_x(Xs, counts, d, chi)
    ctx.set_cores(cores0)
    ctx.sweep_bonds(topts, 2 * (T - 1), restart=True, record=False)          # first sweep: chi grows, flat spectra
    for k in ("svd_calls", "svd_iters_sum", "svd_round2", "svd_jacobi", "svd_fast", "svd_serial"):
        ctx.debug_set(k, 0)
    ctx.profile_enable(True); ctx.profile_reset()
    lo, gn, ch = ctx.sweep_bonds(topts, ns * 2 * (T - 1))
    pr = ctx.profile_get(); ctx.profile_enable(False)
    g = {k: ctx.debug_get(k) for k in ("svd_calls", "svd_iters_sum", "svd_round2", "svd_jacobi", "svd_fast", "svd_serial")}
    print(f"{str(st):24s} svd {pr['svd'][0] / pr['svd'][1]:.3f} ms/split  iters/fast-split {g['svd_iters_sum'] / max(g['svd_fast'], 1):.2f}  "
          f"round2 {g['svd_round2']}  jacobi {g['svd_jacobi']} serial {g['svd_serial']}  of {g['svd_calls']}  final loss {lo[-1]:.6f} mean chi {ch.mean():.1f}", flush=True)
