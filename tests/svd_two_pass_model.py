"""numpy model of the DEVICE algorithm for chi_max > 80 (csrc/svd_subspace.cu: svd_subspace_device + subspace_core):
block subspace iteration with deflating Cholesky-QR, Rayleigh-Ritz through the Gram matrix, NDTensors' truncation rule
with the weight outside the subspace counted as discarded, residual check of the kept pairs, and the deflated second
pass.  Not a restatement of the reference (that is oracle/): it exists so that the algorithm's claims -- same kept
dimension as LAPACK + truncate!, sigma and product to rounding, residuals of pass 2 measured against sigma_1^2 of the
whole matrix -- are checked on the CPU as well (tests/test_host_cpu.py)."""
import numpy as np


def deflating_cholqr(X):
    """X R^-1 with R from the Cholesky factor of X^T X; a pivot below 1e-13 of its column's squared norm deflates the
    column (zero row / column in L and L^-1), as chol_inv_kernel does."""
    G = X.T @ X
    p = G.shape[0]
    A = G.copy()
    d0 = np.diag(G).copy()
    L = np.zeros_like(G)
    for k in range(p):
        dk = A[k, k]
        if dk > 1e-13 * d0[k]:
            L[k:, k] = A[k:, k] / np.sqrt(dk)
            A[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], L[k + 1:, k])
    keep = np.diag(L) > 0
    Linv = np.zeros_like(L)
    if keep.any():
        Linv[np.ix_(keep, keep)] = np.linalg.inv(L[np.ix_(keep, keep)])
    return X @ Linv.T


def truncate(P, trace, maxdim, cutoff, kept=None):
    """ritz_trunc_kernel: P sorted descending; returns the kept count."""
    err = max(trace - (kept or 0.0) - P.sum(), 0.0)
    keep = len(P)
    while keep > maxdim:
        err += P[keep - 1]
        keep -= 1
    floor = 1 if kept is None else 0
    while keep > floor and err + P[keep - 1] <= cutoff * trace:
        err += P[keep - 1]
        keep -= 1
    return max(keep, floor)


def subspace_pass(M, k, p, first, trace, cutoff, rng, kept=None, res_scale=None, log=None):
    """One call of subspace_core: rounds of (first, 3, 3, then contraction-rate sized) iterations until the residual of
    the kept Ritz pairs is <= 5e-14.  Returns (keep, U S, V, sigma^2, iterations) or None (-> exact SVD)."""
    n = M.shape[1]
    Q = rng.uniform(-1.0, 1.0, (n, p))
    its, prev, nxt, budget = 0, None, 3, 48
    for rnd in range(11):
        nit = first if rnd == 0 else nxt
        for _ in range(nit):
            Q = deflating_cholqr(M.T @ deflating_cholqr(M @ Q))
        its += nit
        Q = deflating_cholqr(Q)
        Z = M @ Q
        ev, W = np.linalg.eigh(Z.T @ Z)
        idx = np.argsort(-ev)
        P, W = np.maximum(ev[idx], 0.0), W[:, idx]
        keep = truncate(P, trace, k, cutoff, kept)
        Vk, Uk = Q @ W[:, :k], Z @ W[:, :k]
        T2 = M.T @ Uk
        scale = P[0] if res_scale is None else res_scale
        res = max([np.linalg.norm(T2[:, i] - P[i] * Vk[:, i]) for i in range(keep)], default=0.0) / max(scale, 1e-300)
        if log is not None:
            log.append((rnd, its, keep, res))
        if res <= 5e-14:
            return keep, Uk[:, :keep], Vk[:, :keep], P[:keep], its
        if rnd < 2 and res <= (1e-5 if rnd == 0 else 1e-10):
            prev, nxt = res, 3
            continue
        rate = res ** (1.0 / nit) if (rnd == 0 or not prev) else (res / prev) ** (1.0 / nit)
        need = int(np.ceil(np.log(2e-14 / res) / np.log(rate))) if 0.0 < rate < 0.85 else 1 << 20
        if (rnd == 0 and res > 1e-3) or need > budget:
            return None
        nxt = min(max(need, 3), 16)
        budget -= nxt
        prev = res
    return None


def two_pass_split(M, chi_max, cutoff, rng, K1=64, pmax=112, log=None):
    """svd_subspace_device for k > 80: (kept, U S, V, sigma^2) or None."""
    trace = float((M * M).sum())
    k = min(chi_max, *M.shape)
    r1 = subspace_pass(M, K1, pmax, 7, trace, cutoff, rng, log=log)
    if r1 is None:
        return None
    c1, U1, V1, P1, _ = r1
    if c1 < K1:
        return c1, U1, V1, P1
    M2 = M - U1 @ V1.T
    k2 = k - K1
    p2 = min(max(-(-2 * k2 // 16) * 16, -(-(k2 + 32) // 16) * 16), pmax)
    r2 = subspace_pass(M2, k2, p2, 5 if p2 >= 2 * k2 else 7, trace, cutoff, rng, kept=float(P1.sum()), res_scale=P1[0], log=log)
    if r2 is None:
        return None
    c2, U2, V2, P2, _ = r2
    return c1 + c2, np.hstack([U1, U2]), np.hstack([V1, V2]), np.concatenate([P1, P2])


def lapack_truncated(M, chi_max, cutoff):
    U, s, Vt = np.linalg.svd(M, full_matrices=False)
    P = s * s
    keep = truncate(P, float(P.sum()), chi_max, cutoff)
    return keep, s[:keep], (U[:, :keep] * s[:keep]) @ Vt[:keep]
