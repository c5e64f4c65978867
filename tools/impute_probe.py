"""K8 throughput at the config-D instance shape, series pdf on/off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpstime_jl_b200 as m
T, d, chi, K, n = 256, 16, 64, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
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
outs = {}
for noseries in (1, 0):
    ctx.debug_set("IMPUTE_NOSERIES", noseries)
    ctx.impute_batch(0, X[:, :296], mask[:, :296], grid)
    for method in ("median", "mean"):
        t0 = time.time(); out = ctx.impute_batch(0, X, mask, grid, method=method); dt = time.time() - t0
        outs[(noseries, method)] = out
        print(f"noseries={noseries} {method}: {n / dt:.0f} instances/s", flush=True)
for method in ("median", "mean"):
    a, b = outs[(1, method)], outs[(0, method)]
    print(method, "max |series - direct| =", np.abs(a - b).max(), " differing grid picks:", int((np.abs(a - b) > 1e-9).sum()), "of", int(mask.sum()))
# growth of the mean-method difference along the missing block (random MPS: a chaotic recursion; the first sites must agree)
a, b = outs[(1, "mean")][:, 0, :], outs[(0, "mean")][:, 0, :]
first = mask.T.argmax(axis=1)
for off in (0, 1, 2, 4, 8, 16, 32, 64, 127):
    dv = np.abs(a[np.arange(n), first + off] - b[np.arange(n), first + off])
    print(f"mean: site {off:3d} of the block: max diff {dv.max():.3e}  median diff {np.median(dv):.3e}")
