import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpstime_jl_b200 as m
T, d, chi, K, n = [int(a) for a in sys.argv[1:6]]
dx = float(sys.argv[6])
ctx = m.Context(0)
rng = np.random.default_rng(3)
cores = m.generate_starting_mps(chi, T, d, 1, seed=7)
ctx.model_init(T, 1, d, chi)
ctx.set_cores(cores)
X = np.clip(0.5 * rng.standard_normal((T, n)), -1, 1)
mask = np.zeros((T, n), dtype=np.uint8)
for i, s0 in enumerate(rng.integers(0, T - K + 1, n)):
    mask[s0:s0 + K, i] = 1
grid = m.make_grid((-1.0, 1.0), dx)
ctx.impute_batch(0, X[:, :256], mask[:, :256], grid, method="median")   # warm-up
t0 = time.time()
out = ctx.impute_batch(0, X, mask, grid, method="median")
print("ok", out.shape, time.time() - t0, np.isfinite(out).all())
