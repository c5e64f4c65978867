import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import mpstime_oracle as o
import mpstime_jl_b200 as m
ctx = m.Context(0)
rng = np.random.default_rng(0)
def case(kind, d, chi, C=2):
    if kind == "random":
        B = rng.standard_normal((d * chi * d * chi, C))
    else:  # low rank + graded noise, like a trained bond after a TSGO step
        a = rng.standard_normal((chi, d, chi, C)); b = rng.standard_normal((chi, d, chi))
        B = np.einsum("asmc,mtb->btasc", a, b).reshape(-1, C); B /= np.linalg.norm(B)
        G = rng.standard_normal(B.shape) * np.exp(-rng.uniform(0, 25, size=(B.shape[0], 1)))
        B = B - 0.01 * G / np.linalg.norm(G)
    return B / np.linalg.norm(B)
for kind, d, chi in (("random", 12, 40), ("lowrank", 12, 40), ("random", 16, 64), ("lowrank", 16, 64), ("lowrank", 10, 20), ("lowrank", 4, 12)):
    B = case(kind, d, chi)
    for gl in (True, False):
        for rep in range(2):
            t = time.time(); cl, cr, s = ctx.bond_split(B, d, chi, chi, gl, chi); dt = time.time() - t
        r_l, r_r, rs = o.decompose_bt(B, (chi, d, chi), gl, chi, 1e-10)
        ein = "asmc,mtb->btasc" if gl else "asm,mtbc->btasc"
        print(f"{kind} d={d} chi={chi} left={gl}: wall {dt*1e3:.1f} ms chi {len(s)}/{len(rs)} sigma {np.abs(s-rs).max()/rs.max():.1e} "
              f"prod {np.abs(np.einsum(ein, cl, cr)-np.einsum(ein, r_l, r_r)).max():.1e}", flush=True)
