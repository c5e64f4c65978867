"""K6 (environment update / forward factor) at a bond shape: TFLOP/s of the slab kernel (MI = 4, 2) and of the
shared-memory tile kernel.   python tools/krao_probe.py d chi N"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import mpstime_jl_b200 as m
import mpstime_oracle as o
d, chi, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
T, C = 6, 2
rng = np.random.default_rng(1)
X = rng.uniform(-1, 1, (T, N))
cores = o.random_start_mps(T, d, chi, C, seed=3)
counts = np.array([N // 2, N - N // 2])
ctx = m.Context(0)
ctx.train_load_x(X, counts, d, chi)
ctx.set_cores(cores)
ref = None
for name, fl in (("slab", {}), ("slab MI=4", {"KRAO_SLAB_MI": 4}), ("slab 2x2", {"KRAO_SLAB_MI": 22}), ("reg/tiles", {"KRAO_NOSLAB": 1})):
    for k in ("KRAO_SLAB_MI", "KRAO_NOSLAB"):
        ctx.debug_set(k, fl.get(k, 0))
    ctx.build_env(True)
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(3):
        ctx.build_env(True)
    ms, n, fl_ = ctx.profile_get()["env"]; ctx.profile_enable(False)
    flops = 2.0 * N * (d * chi + (T - 2) * d * chi * chi)
    print(f"{name:10s}: env chain {ms / 3:.3f} ms per build ({n // 3} launches)  {flops / (ms / 3 * 1e-3) / 1e12:6.2f} TFLOP/s", flush=True)
