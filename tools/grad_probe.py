"""Gradient kernel (K2-grad) TFLOP/s for forced stage-size x ring-depth variants of bond_grad_kr_kernel and for the
shared-memory-tile kernel, at a bond shape.   python tools/grad_probe.py d chi N [variants...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import mpstime_jl_b200 as m
import mpstime_oracle as o
d, chi, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
variants = [int(v) for v in sys.argv[4:]] or [0, 643, 644, 324, 326, 164]
rng = np.random.default_rng(1)
counts = np.array([N // 2, N - N // 2])
xl = o.legendre_encode(rng.uniform(-1, 1, N), d); xr = o.legendre_encode(rng.uniform(-1, 1, N), d)
L = rng.standard_normal((N, chi)) / np.sqrt(chi); R = rng.standard_normal((N, chi)) / np.sqrt(chi)
B = rng.standard_normal(((d * chi) ** 2, 2)); B /= np.linalg.norm(B)
ctx = m.Context(0)
ref = None
phases = [int(p) for p in os.environ.get("PHASES", "").split(",") if p] or [None]
for v, ph in [(v, ph) for v in variants + [-1] for ph in phases]:
    if ph is not None:
        ctx.debug_set("GRAD_PHASES", ph)
    ctx.debug_set("GRAD_NOKR", 1 if v < 0 else 0)
    ctx.debug_set("GRAD_KC", max(v, 0))
    try:
        ctx.bond_loss_grad(B, L, R, xl, xr, counts)
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(3):
            lo, G = ctx.bond_loss_grad(B, L, R, xl, xr, counts)
        ms, n, fl = ctx.profile_get()["grad_kernel"]; ctx.profile_enable(False)
        if ref is None:
            ref = G
        print(f"variant {v:4d} phases {ph}: kernel {'kr' if ctx.debug_get('grad_kernel') == 1 else 'tiles'} {ctx.debug_get('grad_variant')}  "
              f"{fl / (ms * 1e-3) / 1e12:6.2f} TFLOP/s  ({ms / n:.3f} ms)  max|G - G0| {np.abs(G - ref).max():.2e}", flush=True)
    except Exception as e:
        print(f"variant {v}: {e}")
