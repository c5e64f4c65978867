"""A/B of the K5 building blocks on ONE split at a bond shape, device time from the library's own CUDA events
(profile slot 'svd'), fixed iteration count so every variant does the same work:
    python tools/svd_ab.py [d chi C iters]
variants: overlapped loop / serial loop, blocked / unblocked Cholesky inverse, register / shared-memory Jacobi."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpstime_jl_b200 as m
d, chi, C, its = (int(a) for a in (sys.argv[1:5] + [16, 64, 2, 6][len(sys.argv) - 1:]))
rng = np.random.default_rng(0)
D = d * chi
U, _ = np.linalg.qr(rng.standard_normal((C * D, 2 * chi)))
V, _ = np.linalg.qr(rng.standard_normal((D, 2 * chi)))
s = np.concatenate([np.logspace(0, -2.5, chi), 1e-3 * np.logspace(0, -2, chi)])
M = (U * s) @ V.T + 1e-5 * rng.standard_normal((C * D, D)) / np.sqrt(D)
M /= np.linalg.norm(M)
B = np.ascontiguousarray(M.reshape(C, D, D).transpose(0, 2, 1).reshape(C, D * D).T)
sig_ref = np.linalg.svd(M, compute_uv=False)[:chi]
ctx = m.Context(0)
ctx.debug_set("SVD_IT", its)
VARIANTS = [("default", {}), ("no graph", {"SVD_NOGRAPH": 1}),
            ("idle start", {"SVD_SYNCFIRST": 1}), ("idle start, no upload", {"SVD_SYNCFIRST": 1, "SVD_NOPREP": 1}),
            ("idle start, no graph", {"SVD_SYNCFIRST": 1, "SVD_NOGRAPH": 1}), ("serial loop", {"SVD_SERIAL": 1}), ("unblocked chol", {"SVD_CHOLSEQ": 1}),
            ("serial + unblocked chol", {"SVD_SERIAL": 1, "SVD_CHOLSEQ": 1}), ("smem jacobi", {"SVD_EIGSMEM": 1}),
            ("round-1 configuration", {"SVD_SERIAL": 1, "SVD_CHOLSEQ": 1, "SVD_EIGSMEM": 1})]
ref = None
for name, fl in VARIANTS:
    for k in ("SVD_SERIAL", "SVD_CHOLSEQ", "SVD_EIGSMEM", "SVD_NOGRAPH", "SVD_SYNCFIRST", "SVD_NOPREP"):
        ctx.debug_set(k, fl.get(k, 0))
    try:
        res = ctx.bond_split(B, d, chi, chi, True, chi)
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(8):
            res = ctx.bond_split(B, d, chi, chi, True, chi)
        ms, n, _w = ctx.profile_get()["svd"]; ctx.profile_enable(False)
        prod = np.einsum("askc,ktb->astbc", res[0], res[1])
        if ref is None:
            ref = prod
        sg = res[2]
        print(f"{name:26s} {ms / n:7.3f} ms/split  path {ctx.debug_get('svd_path')} iters {ctx.debug_get('svd_iters')} chi {len(sg)}  "
              f"sigma vs LAPACK {np.abs(sg - sig_ref[:len(sg)]).max() / sig_ref[0]:.1e}  product vs default {np.abs(prod - ref).max():.1e}", flush=True)
    except Exception as e:
        print(name, "FAILED", e, flush=True)
