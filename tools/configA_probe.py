#!/usr/bin/env python
"""Where the time of the small reference case goes (BASELINE.json configs[0] shape: 67 train / 1029 test, T = 24,
d = 10, chi_max = 20, 5 sweeps): wall clock of fitMPS, device time per kernel family from the library's events, SVD path
statistics; with and without the CUDA graphs of the subspace rounds."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpstime_jl_b200 as m  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    T = 24

    def ipd_like(n):
        y = rng.integers(0, 2, n)
        t = np.arange(T)
        base = np.where(y[:, None] == 0, np.sin(2 * np.pi * t / 24.0)[None, :], np.sin(2 * np.pi * t / 24.0 + 0.6)[None, :] * 0.8)
        return base + 0.25 * rng.standard_normal((n, T)), y
    Xtr, ytr = ipd_like(67)
    Xte, yte = ipd_like(1029)
    out = []
    for log_level in (3, 0):
        opts = m.MPSOptions(d=10, chi_max=20, nsweeps=5, verbosity=-1, log_level=log_level)
        m.fitMPS(Xtr, ytr, Xte, yte, opts)
        ctx = m.api._context()
        for nograph in (0, 1):
            ctx.debug_set("SVD_NOGRAPH", nograph)
            keys = ("svd_calls", "svd_iters_sum", "svd_round2", "svd_jacobi", "svd_fast", "svd_serial")
            for k in keys:
                ctx.debug_set(k, 0)
            ctx.profile_enable(True)
            ctx.profile_reset()
            t0 = time.time()
            m.fitMPS(Xtr, ytr, Xte, yte, opts)
            wall = time.time() - t0
            pr = ctx.profile_get()
            ctx.profile_enable(False)
            row = {"log_level": log_level, "nograph": nograph, "fitMPS_seconds": wall,
                   "device_ms": {k: round(v[0], 2) for k, v in pr.items()}, "launches": {k: int(v[1]) for k, v in pr.items()},
                   "svd": {k: ctx.debug_get(k) for k in keys}}
            out.append(row)
            print(json.dumps(row), flush=True)
        ctx.debug_set("SVD_NOGRAPH", 0)
    # host-only share: normalisation + sort + table checks, no device work
    t0 = time.time()
    Xs, norms = m.transform_train_data(Xtr.T, opts)
    m.transform_test_data(Xte.T, norms, opts)
    print(json.dumps({"host_transform_seconds": time.time() - t0}))
    json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_configA_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
