"""K8 at the config-D instance shape on a small batch (2 instances per SM): the kernel's own cycle breakdown
(MPST_IMPUTE_DEBUG=1) or a target for `ncu --set full -k regex:impute_kernel`.  python tools/impute_prof.py [n] [method]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpstime_jl_b200 as m
T, d, chi, K = 256, 16, 64, 128
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
method = sys.argv[2] if len(sys.argv) > 2 else "median"
rng = np.random.default_rng(3)
cores = m.generate_starting_mps(chi, T, d, 1, seed=7)
ctx = m.Context(0)
ctx.model_init(T, 1, d, chi); ctx.set_cores(cores)
t = np.arange(1, T + 1)
X = np.clip(np.sin(2 * np.pi * t[:, None] / 24.0 + rng.uniform(0, 2 * np.pi, n)[None, :]) * 0.45 + 0.04 * rng.standard_normal((T, n)), -1, 1)
mask = np.zeros((T, n), dtype=np.uint8)
for i, s0 in enumerate(rng.integers(0, T - K + 1, n)):
    mask[s0:s0 + K, i] = 1
grid = m.make_grid((-1.0, 1.0), 1e-4)
U = rng.uniform(0.0, 1.0, size=(n, 1, K))
t0 = time.time()
out = ctx.impute_batch(0, X, mask, grid, method=method, uniforms=U if method == "ITS" else None)
print(f"{method}: {n} instances in {time.time() - t0:.3f} s", flush=True)
