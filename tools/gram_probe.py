#!/usr/bin/env python
"""Device time of the Gram-path splits (d * chi <= 112 columns): Cholesky + register-resident Jacobi (opt-in, MPST_SVD_GRAMREG) against
the default shared-memory Jacobi on the columns of H."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpstime_jl_b200 as m  # noqa: E402

ctx = m.Context(0)
rng = np.random.default_rng(1)
rows = []
for d, chi in ((6, 16), (10, 10), (5, 20), (2, 15), (4, 8)):
    C, n = 2, d * chi
    U, _ = np.linalg.qr(rng.standard_normal((C * n, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    M = (U * 0.9 ** np.arange(n)) @ V.T
    M /= np.linalg.norm(M)
    B = np.ascontiguousarray(M.reshape(chi, C, d, d, chi).transpose(1, 4, 3, 0, 2).reshape(C, -1).T)
    row = {"d": d, "chi": chi, "m": C * n, "n": n}
    outs = {}
    for name, flag in (("chol_reg", 0), ("smem", 1)):
        ctx.debug_set("SVD_GRAMREG", 1 - flag)
        ctx.bond_split(B, d, chi, chi, True, chi)
        ctx.profile_enable(True)
        ctx.profile_reset()
        for _ in range(5):
            outs[name] = ctx.bond_split(B, d, chi, chi, True, chi)
        row[name + "_svd_ms_device"] = ctx.profile_get()["svd"][0] / 5
        row[name + "_path"] = ctx.debug_get("svd_path")
        ctx.profile_enable(False)
    ctx.debug_set("SVD_GRAMREG", 0)
    a, b = outs["chol_reg"], outs["smem"]
    row["kept"] = (len(a[2]), len(b[2]))
    row["sigma_diff"] = float(np.abs(a[2] - b[2]).max()) if len(a[2]) == len(b[2]) else None
    ein = "asmc,mtb->btasc"
    row["product_diff"] = float(np.abs(np.einsum(ein, a[0], a[1]) - np.einsum(ein, b[0], b[1])).max()) if len(a[2]) == len(b[2]) else None
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_gram_probe.json"), "w"), indent=1)
