"""A/B of the Rayleigh-Ritz eigen-solvers inside the subspace SVD (register-resident vs shared-memory Jacobi):
same split, both solvers, sigma / product agreement and wall time per split.  python tools/eig_ab.py [d chi C]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpstime_jl_b200 as m
d, chi, C = (int(a) for a in (sys.argv[1:4] + [16, 64, 2][len(sys.argv) - 1:]))
rng = np.random.default_rng(0)
D = d * chi
# decaying spectrum: rank-chi part + noise at 1e-2 (a trained bond after a gradient step)
U, _ = np.linalg.qr(rng.standard_normal((C * D, 2 * chi)))
V, _ = np.linalg.qr(rng.standard_normal((D, 2 * chi)))
s = np.concatenate([np.logspace(0, -2.5, chi), 1e-3 * np.logspace(0, -2, chi)])
M = (U * s) @ V.T + 1e-5 * rng.standard_normal((C * D, D)) / np.sqrt(D)
M /= np.linalg.norm(M)
# BondTensor layout [C][p + Dl*q] (going_left: rows (c, p) x cols q)
B = np.ascontiguousarray(M.reshape(C, D, D).transpose(0, 2, 1).reshape(C, D * D).T)   # (D*D, C), column = class
ctx = m.Context(0)
sig_ref = np.linalg.svd(M, compute_uv=False)[:chi]
out = {}
for name, flag in (("reg", 0), ("smem", 1)):
    ctx.debug_set("SVD_EIGSMEM", flag)
    res = ctx.bond_split(B, d, chi, chi, True, chi)
    ts = []
    for _ in range(5):
        t = time.time(); res = ctx.bond_split(B, d, chi, chi, True, chi); ts.append(time.time() - t)
    sig = np.asarray(res[-1] if isinstance(res, tuple) else res)
    out[name] = res
    print(name, "svd_path", ctx.debug_get("svd_path"), "iters", ctx.debug_get("svd_iters"), "min wall ms %.3f" % (1e3 * min(ts)), flush=True)
def prod(res):
    cl, cr, sg = res
    return np.einsum("askc,ktb->astbc", cl, cr)
pr, ps = prod(out["reg"]), prod(out["smem"])
print("chi", len(out["reg"][2]), len(out["smem"][2]), "product max|reg - smem| %.3e" % np.abs(pr - ps).max(),
      "sigma max|reg - smem| %.3e" % np.abs(out["reg"][2] - out["smem"][2]).max())
sig = out["reg"][2]
print("sigma vs LAPACK rel %.3e" % (np.abs(sig - sig_ref[:len(sig)]) / sig_ref[0]).max())
cr = out["reg"][1].reshape(len(sig), -1)
print("ortho core orthonormality %.3e" % np.abs(cr @ cr.T - np.eye(len(sig))).max())
