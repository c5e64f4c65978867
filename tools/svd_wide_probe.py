#!/usr/bin/env python
"""Time of one truncated split at chi_max = 128 (configs[4] shapes): deflated two-pass subspace iteration against the
exact Jacobi it replaces (MPST_SVD_NO2PASS).  Host wall time of mpst_bond_split (includes the H2D copy of the bond
tensor and the D2H copy of both cores, identical on both arms); best of 5 after one warm call."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpstime_jl_b200 as m  # noqa: E402


def decaying(rng, mrows, n, r):
    U, _ = np.linalg.qr(rng.standard_normal((mrows, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    M = (U * r ** np.arange(n)) @ V.T
    return M / np.linalg.norm(M)


def main():
    ctx = m.Context(0)
    rows = []
    for d, chi, r in ((6, 128, 0.95), (12, 128, 0.97), (24, 128, 0.985)):
        rng = np.random.default_rng(d)
        C = 2
        M = decaying(rng, chi * C * d, d * chi, r)
        B = np.ascontiguousarray(M.reshape(chi, C, d, d, chi).transpose(1, 4, 3, 0, 2).reshape(C, -1).T)
        row = {"d": d, "chi": chi, "m": chi * C * d, "n": d * chi}
        for name, flag in (("two_pass", 0), ("jacobi", 1)):
            ctx.debug_set("SVD_NO2PASS", flag)
            ctx.bond_split(B, d, chi, chi, True, chi)
            best = 1e9
            for _ in range(5 if flag == 0 else 2):
                t0 = time.perf_counter()
                out = ctx.bond_split(B, d, chi, chi, True, chi)
                best = min(best, time.perf_counter() - t0)
            row[name + "_ms_host_call"] = 1e3 * best
            row[name + "_path"] = ctx.debug_get("svd_path")
            row[name + "_chi"] = len(out[2])
            row[name + "_iters"] = ctx.debug_get("svd_iters")
        ctx.debug_set("SVD_NO2PASS", 0)
        # the copies alone (same call with a trivially small chi_max would change the path): time H2D + D2H separately
        rows.append(row)
        print(json.dumps(row), flush=True)
    json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_svd_wide_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
