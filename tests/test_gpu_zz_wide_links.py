"""chi_max > 80: the truncated split as a deflated TWO-PASS subspace iteration (svd_subspace.cu, `svd_subspace_device`).
BASELINE.json configs[4] (the chi x d micro-sweep) goes to chi = 128; the single-CTA Cholesky / Rayleigh-Ritz kernels
stop at 128 columns, so these bonds used to take the exact Jacobi SVD (tens of ms).  Same bar as the single-pass tests
(tests/test_gpu_bench_shapes.py): kept dimension, singular values and the truncated two-site product equal LAPACK +
NDTensors' truncate! as restated by the oracle (reference: src/Training/RealRealHighDimension.jl:146-203), and the test
asserts that the fast path -- not the Jacobi fallback -- produced them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SVD_SUBSPACE = 3
EIN = "asmc,mtb->btasc"


def _bond_from_matrix(M, d, cl, cr, C):
    # going left: rows (a, c, s_l), columns (s_r, b) -> (D, C) bond tensor, index s_l + d*(a + cl*(s_r + d*b))
    return np.ascontiguousarray(M.reshape(cl, C, d, d, cr).transpose(1, 4, 3, 0, 2).reshape(C, -1).T)


def _decaying_matrix(rng, m, n, knee, r1, r2):
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    k = np.arange(n)
    s = np.where(k < knee, r1 ** np.minimum(k, knee), r1 ** knee * r2 ** np.maximum(k - knee, 0))
    M = (U * s) @ V.T
    return M / np.linalg.norm(M)


def _check(ctx, oracle, B, d, chi, chimax, want_twopass):
    ctx.debug_set("svd_twopass", 0)
    c_l, c_r, sig = ctx.bond_split(B, d, chi, chi, True, chimax)
    path, two = ctx.debug_get("svd_path"), ctx.debug_get("svd_twopass")
    r_l, r_r, rs = oracle.decompose_bt(B, (chi, d, chi), True, chimax, 1e-10)
    assert len(sig) == len(rs), (len(sig), len(rs))
    assert np.abs(sig - rs).max() < 1e-10 * rs.max()
    assert np.abs(np.einsum(EIN, c_l, c_r) - np.einsum(EIN, r_l, r_r)).max() < 1e-10
    # the orthogonal core is right-canonical: V^T V = 1 (V2 against V1 only to eps sigma_1 / sigma_j, see DESIGN)
    V = c_r.reshape(len(sig), -1)
    assert np.abs(V @ V.T - np.eye(len(sig))).max() < 1e-9
    assert path == SVD_SUBSPACE, f"took path {path}"
    assert two == (1 if want_twopass else 0), two
    return len(rs)


@pytest.mark.parametrize("d,chi,knee,r1,r2,chimax,kept_range,twopass", [
    (6, 128, 0, 0.95, 0.95, 128, (128, 128), True),     # 1536 x 768, chi_max decides, 64 + 64
    (6, 128, 0, 0.90, 0.90, 128, (100, 120), True),     # the 1e-10 cutoff stops inside the second pass
    (6, 128, 30, 0.6, 0.99, 128, (16, 40), False),      # cutoff stops inside the first pass: no second pass
    (6, 128, 0, 0.93, 0.93, 96, (96, 96), True),        # chi_max = 96: 64 + 32
    (12, 128, 0, 0.97, 0.97, 128, (128, 128), True),    # 3072 x 1536 (d = 12, chi = 128)
])
def test_two_pass_split_equals_lapack(ctx, oracle, d, chi, knee, r1, r2, chimax, kept_range, twopass):
    rng = np.random.default_rng(int(1000 * r1) + d + chimax)
    C = 2
    M = _decaying_matrix(rng, chi * C * d, d * chi, knee, r1, r2)
    kept = _check(ctx, oracle, _bond_from_matrix(M, d, chi, chi, C), d, chi, chimax, twopass)
    assert kept_range[0] <= kept <= kept_range[1], kept


@pytest.mark.parametrize("rank", [64, 70, 100])
def test_two_pass_split_rank_deficient(ctx, oracle, rank):
    """Exact rank 64: the second pass sees rounding noise only and keeps nothing; rank 70 / 100: it keeps the rest."""
    rng = np.random.default_rng(rank)
    d, chi, C = 6, 128, 2
    m, n = chi * C * d, d * chi
    U, _ = np.linalg.qr(rng.standard_normal((m, rank)))
    V, _ = np.linalg.qr(rng.standard_normal((n, rank)))
    M = (U * (0.9 ** np.arange(rank))) @ V.T
    M /= np.linalg.norm(M)
    kept = _check(ctx, oracle, _bond_from_matrix(M, d, chi, chi, C), d, chi, chi, True)
    assert kept == rank


def test_two_pass_switch_off_takes_jacobi(ctx, oracle):
    """MPST_SVD_NO2PASS restores the previous behaviour (exact Jacobi above chi_max = 80) with the same result."""
    rng = np.random.default_rng(3)
    d, chi, C = 6, 128, 2
    B = _bond_from_matrix(_decaying_matrix(rng, chi * C * d, d * chi, 0, 0.95, 0.95), d, chi, chi, C)
    a = ctx.bond_split(B, d, chi, chi, True, chi)
    assert ctx.debug_get("svd_path") == SVD_SUBSPACE
    try:
        ctx.debug_set("SVD_NO2PASS", 1)
        b = ctx.bond_split(B, d, chi, chi, True, chi)
        assert ctx.debug_get("svd_path") in (4, 5)
    finally:
        ctx.debug_set("SVD_NO2PASS", 0)
    assert len(a[2]) == len(b[2]) and np.abs(a[2] - b[2]).max() < 1e-12
    assert np.abs(np.einsum(EIN, a[0], a[1]) - np.einsum(EIN, b[0], b[1])).max() < 1e-12


@pytest.mark.parametrize("d,chi,r,min_iters", [(12, 40, 0.985, 12), (16, 64, 0.985, 12)])
def test_slow_decay_keeps_iterating_instead_of_jacobi(ctx, oracle, d, chi, r, min_iters):
    """A slowly decaying spectrum (contraction ~0.25 per iteration) misses the residual bound after the standard rounds.
    The observed contraction rate predicts convergence within the iteration budget, so the split stays on the subspace
    path (it used to take the exact Jacobi: 20 ms at 960 x 480, seconds at 6144 x 3072)."""
    rng = np.random.default_rng(int(1000 * r) + chi)
    C = 2
    M = _decaying_matrix(rng, chi * C * d, d * chi, 0, r, r)
    kept = _check(ctx, oracle, _bond_from_matrix(M, d, chi, chi, C), d, chi, chi, False)
    assert kept == chi
    assert ctx.debug_get("svd_iters") >= min_iters, ctx.debug_get("svd_iters")


def test_teacher_forced_bonds_with_wide_links(ctx, oracle, pkg):
    """chi_max = 96 inside a sweep (mpst_bond_step, not the stand-alone split): bonds of a device-trained chain whose
    links are saturated at 96 take the two-pass split; loss, gradient norm, kept dimension, singular values and the new
    two-site product equal the oracle's bond from the same state."""
    N, T, d, chi, eta = 512, 10, 6, 96, 0.05
    X, y = oracle.synthetic_two_class(N, T, seed=4)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    ctx.train_load_x(Xs[:, order], counts, d, chi)
    ctx.set_cores(oracle.random_start_mps(T, d, 4, 2, seed=3))
    opts = pkg.make_opts(chi_max=chi, eta=eta)
    ctx.sweep_bonds(opts, 2 * 2 * (T - 1), restart=True, record=False)
    seen = 0
    for j in range(T - 2, 1, -1):
        cs = ctx.get_cores()
        saturated = cs[j].shape[0] == chi and cs[j + 1].shape[2] == chi
        ctx.debug_set("svd_twopass", 0)
        lo, gn, k = ctx.bond_step(j, True, opts)
        if not saturated:
            continue
        path, two = ctx.debug_get("svd_path"), ctx.debug_get("svd_twopass")
        L = np.ones((N, 1))
        for q in range(j):
            L = oracle.env_step_left(phi[:, q], L, cs[q])
        R = np.ones((N, 1))
        for q in range(T - 1, j + 1, -1):
            R = oracle.env_step_right(phi[:, q], R, cs[q])
        B, dims = oracle.flatten_bt(cs[j], cs[j + 1])
        Bn, lo_r, gn_r = oracle.apply_update(B, L, R, phi[:, j], phi[:, j + 1], counts, eta=eta)
        cl, cr, S = oracle.decompose_bt(Bn, dims, True, chi, 1e-10)
        assert abs(lo - lo_r) <= 1e-9 * abs(lo_r) and abs(gn - gn_r) <= 1e-9 * gn_r
        assert k == len(S), (j, k, len(S))
        dl, dr = ctx.get_core(j), ctx.get_core(j + 1)
        assert np.abs(np.einsum(EIN, dl, dr) - np.einsum(EIN, cl, cr)).max() < 1e-10
        sig_d = np.linalg.svd(dl.transpose(0, 3, 1, 2).reshape(-1, k), compute_uv=False)
        assert np.abs(sig_d - S).max() < 1e-10 * S.max()
        assert path == SVD_SUBSPACE and two == 1, (j, path, two)
        seen += 1
    assert seen >= 1, "no chi-saturated bond was compared"
