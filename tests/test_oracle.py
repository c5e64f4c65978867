"""CPU tests of the oracle itself: pinned against the reference's own serialized output where that
exists (tests/golden/ecg200_legendre.npz, made by tests/golden/make_golden_from_jld2.py), otherwise
self-consistency (literal loop form == vectorised form, gradient == finite differences, ...)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ecg200_legendre.npz")


def test_golden_normalisation_and_legendre(oracle):
    """The reference's saved PStates (ECG200, Legendre_No_Norm d=5, sigmoid+minmax) are reproduced from
    its saved raw matrix: pins transform_train_data (utils.jl:161-200) and legendre_encode
    (bases.jl:77-92) to 1e-14."""
    g = np.load(GOLD)
    Xs, norms = oracle.transform_train_data(g["X_orig"].T)
    assert Xs.min() == -1.0 and Xs.max() == 1.0
    phi = oracle.encode(Xs.T, 5, "legendre_no_norm")
    assert np.abs(phi - g["phi_ref"]).max() < 1e-14


def test_legendre_orthonormal(oracle):
    x, w = np.polynomial.legendre.leggauss(64)
    phi = oracle.legendre_encode(x, 16)
    gram = (phi * w[:, None]).T @ phi
    assert np.abs(gram - np.eye(16)).max() < 1e-12
    assert np.allclose(oracle.legendre_encode(x, 6, norm=True), oracle.legendre_encode(x, 6) / np.sqrt(np.sqrt(6.5) * 6))


def test_fourier_sahand_stoudenmire(oracle):
    assert list(oracle.fourier_freqs(6)) == [0, 1, -1, 2, -2, 3]
    x = np.linspace(0, 1, 11)
    f = oracle.fourier_encode(2 * x - 1, 5)
    assert np.allclose(np.sum(np.abs(f) ** 2, axis=-1), 1.0)
    s = oracle.stoudenmire_encode(x)
    assert np.allclose(np.sum(np.abs(s) ** 2, axis=-1), 1.0)
    sa = oracle.sahand_encode(x, 4)
    assert sa.shape == (11, 4) and np.all(np.abs(sa[0, 2:]) == 0)


def test_transform_roundtrip_and_oob(oracle):
    rng = np.random.default_rng(0)
    Xtr = rng.standard_normal((30, 40))
    Xs, norms = oracle.transform_train_data(Xtr)
    assert Xs.min() >= -1 and Xs.max() <= 1
    Xte = 1.5 * rng.standard_normal((30, 7))
    Xt, oob = oracle.transform_test_data(Xte, norms)
    assert Xt.min() >= -1 - 1e-12 and Xt.max() <= 1 + 1e-12 and len(oob) > 0
    back = oracle.invert_test_transform(Xt, oob, norms)
    ok = np.isfinite(back)
    assert ok.mean() > 0.9 and np.abs(back[ok] - Xte[ok]).max() < 1e-8


def _problem(oracle, N=40, T=8, d=4, C=2, seed=1):
    X, y = oracle.synthetic_two_class(N, T, seed=seed)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.random_start_mps(T, d, 4, C, seed=3)
    return phi, ys, counts, cores


def test_loss_grad_loop_vs_vectorised_and_fd(oracle):
    phi, ys, counts, cores = _problem(oracle)
    N, T, d = phi.shape
    LE = oracle.construct_caches(cores, phi, True)
    l, r = T - 2, T - 1
    B, dims = oracle.flatten_bt(cores[l], cores[r])
    L, R = LE[l - 1], np.ones((N, 1))
    xl, xr = phi[:, l], phi[:, r]
    for sep in (False, True):
        a = oracle.loss_grad_KLD_loop(B, L, R, xl, xr, counts, sep)
        b = oracle.loss_grad_KLD(B, L, R, xl, xr, counts, sep)
        assert abs(a[0] - b[0]) < 1e-12 and np.abs(a[1] - b[1]).max() < 1e-10
    a = oracle.loss_grad_MSE_loop(B, L, R, xl, xr, counts)
    b = oracle.loss_grad_MSE(B, L, R, xl, xr, counts)
    assert abs(a[0] - b[0]) < 1e-13 and np.abs(a[1] - b[1]).max() < 1e-13
    rng = np.random.default_rng(0)
    E = rng.standard_normal(B.shape)
    eps = 1e-6
    lo, G = oracle.loss_grad_KLD(B, L, R, xl, xr, counts)
    fd = (oracle.loss_grad_KLD(B + eps * E, L, R, xl, xr, counts)[0] - oracle.loss_grad_KLD(B - eps * E, L, R, xl, xr, counts)[0]) / (2 * eps)
    assert abs(fd - 2 * np.sum(G * E)) < 1e-5 * abs(fd)      # the reference's KLD gradient omits the factor 2
    lo, G = oracle.loss_grad_MSE(B, L, R, xl, xr, counts)
    fd = (oracle.loss_grad_MSE(B + eps * E, L, R, xl, xr, counts)[0] - oracle.loss_grad_MSE(B - eps * E, L, R, xl, xr, counts)[0]) / (2 * eps)
    assert abs(fd - np.sum(G * E)) < 1e-6 * max(1.0, abs(fd))
    # bond overlap == full chain contraction, KLD loss == KL_div of the chain (summary.jl:459-471)
    yh = oracle.bond_yhat(B, L, R, xl, xr)
    assert np.abs(yh - oracle.overlaps(cores, phi)).max() < 1e-13
    mse, kld, acc = oracle.mse_loss_acc(cores, phi, ys)
    assert abs(kld - oracle.loss_grad_KLD(B, L, R, xl, xr, counts)[0]) < 1e-12


def test_truncation_rule(oracle):
    P = np.array([1.0, 0.5, 1e-3, 1e-12, 1e-13, 0.0])
    assert oracle.truncate_spectrum(P, 10, 1e-10) == 3
    assert oracle.truncate_spectrum(P, 2, 1e-10) == 2
    assert oracle.truncate_spectrum(P, 10, 0.0) == 5            # exact zeros satisfy err + 0 <= 0
    assert oracle.truncate_spectrum(np.array([0.0, 0.0]), 5, 1e-10) == 1
    # weight already discarded by maxdim counts towards the cutoff sum
    P = np.array([1.0, 1e-3, 6e-11, 6e-11])
    assert oracle.truncate_spectrum(P, 3, 1e-10) == 3
    assert oracle.truncate_spectrum(P[:3], 3, 1e-10) == 2


def test_decompose_reconstructs(oracle):
    rng = np.random.default_rng(1)
    d, cl, cr, C = 3, 4, 5, 2
    B = rng.standard_normal((d * cl * d * cr, C))
    for gl in (True, False):
        a, b, S = oracle.decompose_bt(B, (cl, d, cr), gl, 100, 0.0)
        full = np.einsum("asmc,mtb->btasc", a, b) if gl else np.einsum("asm,mtbc->btasc", a, b)
        assert np.abs(full.reshape(-1, C) - B).max() < 1e-12
        a, b, S = oracle.decompose_bt(B, (cl, d, cr), gl, 5, 1e-10)
        assert len(S) == 5


def test_sweep_decreases_loss_and_loop_equals_vectorised(oracle):
    phi, ys, counts, cores = _problem(oracle)
    rec, rec2 = [], []
    new = oracle.fit_sweeps(cores, phi, counts, nsweeps=3, chi_max=10, eta=0.05, record=rec)
    per = [np.mean([r["loss"] for r in rec if r["sweep"] == s]) for s in range(3)]
    assert per[0] > per[1] > per[2]
    assert abs(oracle._norm2_general(new) - 1.0) < 1e-12
    oracle.fit_sweeps(cores, phi, counts, nsweeps=1, chi_max=10, eta=0.05, record=rec2, loop=True)
    assert max(abs(a["loss"] - b["loss"]) for a, b in zip(rec, rec2)) < 1e-9
    assert oracle.mse_loss_acc(new, phi, ys)[2] >= oracle.mse_loss_acc(cores, phi, ys)[2]


def test_imputation_gram_equals_qr(oracle):
    """orthogonalize!-based rho (MPS_methods.jl:110,152) == right-Gram formulation."""
    phi, ys, counts, cores = _problem(oracle, N=60, T=8)
    new = oracle.fit_sweeps(cores, phi, counts, nsweeps=2, chi_max=8, eta=0.05)
    cls = oracle.expand_label_index(new)[0]
    assert abs(oracle._norm2_general(cls) - 1.0) < 1e-12
    d = phi.shape[2]
    x = np.linspace(-0.5, 0.5, 8)
    missing = [2, 3, 5]
    cond = oracle.precondition(cls, oracle.encode(x, d), missing)
    orth = oracle.orthogonalize_to_first(cond)
    A = orth[0][0]
    rho_qr = A @ A.T
    G = np.ones((1, 1))
    for k in range(len(cond) - 1, 0, -1):
        G = np.einsum("asb,bc,xsc->ax", cond[k], G, cond[k])
    rho_gram = np.einsum("sb,bc,tc->st", cond[0][0], G, cond[0][0])
    assert np.abs(rho_qr - rho_gram).max() < 1e-12 * np.abs(rho_qr).max()
    grid = oracle.make_grid()
    assert len(grid) == 20001 and grid[0] == -1.0 and grid[-1] == 1.0
    genc = oracle.encode(grid, d)
    out, idx = oracle.impute_series(cls, x, missing, grid, genc, d)
    assert np.all(out[[0, 1, 4, 6, 7]] == x[[0, 1, 4, 6, 7]]) and np.all(np.abs(out[missing]) <= 1)
    u = [0.5, 0.5, 0.5]
    out2, idx2 = oracle.impute_series(cls, x, missing, grid, genc, d, method="ITS", uniforms=u)
    assert np.array_equal(idx, idx2)                        # ITS at u = 1/2 is the median


def test_c_restatement_of_hot_loop_matches_numpy_oracle(oracle):
    """oracle/bond_ref.c (the timed CPU baseline's hot loop) against the literal numpy loop and the vectorised form:
    KLD loss + gradient (1 and 3 threads, both normalisations) and the environment update."""
    import bond_ref
    rng = np.random.default_rng(5)
    N, d, cl, cr, C = 70, 3, 4, 5, 2
    counts = np.array([30, 40])
    xl = oracle.legendre_encode(rng.uniform(-1, 1, N), d)
    xr = oracle.legendre_encode(rng.uniform(-1, 1, N), d)
    L = rng.standard_normal((N, cl))
    R = rng.standard_normal((N, cr))
    B = rng.standard_normal((d * cl * d * cr, C))
    for sep in (False, True):
        lo_l, G_l = oracle.loss_grad_KLD_loop(B, L, R, xl, xr, counts, sep)
        lo_v, G_v = oracle.loss_grad_KLD(B, L, R, xl, xr, counts, sep)
        for nt in (1, 3):
            lo_c, G_c = bond_ref.loss_grad_kld(B, L, R, xl, xr, counts, sep, nt)
            assert abs(lo_c - lo_l) <= 1e-12 * abs(lo_l) and abs(lo_c - lo_v) <= 1e-12 * abs(lo_v)
            assert np.abs(G_c - G_l).max() <= 1e-11 * np.abs(G_l).max()
            assert np.abs(G_c - G_v).max() <= 1e-11 * np.abs(G_v).max()
    core = rng.standard_normal((cl, d, 6))
    env = bond_ref.env_update(xl, L, core.transpose(1, 0, 2).reshape(-1, order="F"), 6)      # [s + d*(a + chi*k)]
    assert np.abs(env - oracle.env_step_left(xl, L, core)).max() < 1e-12


def test_reference_trained_mps_pins_layout_norm_and_classification(oracle):
    """The reference's OWN trained MPS (ECG200 fixture, extracted by tests/golden/make_golden_mps_from_jld2.py) pins the
    oracle's core index order, label position, normalize! and contract_mps/classify: with the reference's own encoded
    training set (ecg200_legendre.npz) the overlaps must classify all 100 training series correctly (the fixture is a
    converged fit) and <W|W> must be 1 -- any wrong index order or conjugation breaks both."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ecg200_trained_mps.npz"))
    e = np.load(os.path.join(os.path.dirname(__file__), "golden", "ecg200_legendre.npz"))
    T = 96
    cores = [g["core_%02d" % j] for j in range(T)]
    assert int(g["label_pos"]) == T - 1
    assert cores[0].shape[0] == 1 and cores[-1].shape[2] == 1 and cores[-1].shape[3] == 2
    assert max(A.shape[2] for A in cores) == 25                       # the fixture's chi_max
    assert abs(oracle._norm2_general(cores) - 1.0) < 1e-12             # normalize!(W), RealRealHighDimension.jl:852
    y = np.r_[np.zeros(31, dtype=np.int64), np.ones(69, dtype=np.int64)]
    yh = oracle.overlaps(cores, e["phi_ref"])
    assert np.array_equal(np.argmax(yh * yh, axis=1), y)
    assert np.array_equal(oracle.classify(cores, e["phi_ref"]), y)
    # the last forward half-sweep leaves sites 0..T-2 left-orthonormal (U of decomposeBT, :177-196)
    for A in cores[:-1]:
        a, s, b = A.shape
        M = A.reshape(a * s, b)
        assert np.abs(M.T @ M - np.eye(b)).max() < 1e-10
    # ... and the label core carries the singular values: its Gram matrix over (class, site) is diagonal
    A = cores[-1][:, :, 0, :]
    Gm = np.einsum("asc,bsc->ab", A, A)
    assert np.abs(Gm - np.diag(np.diag(Gm))).max() < 1e-10 * np.abs(Gm).max()
    sig = np.sqrt(np.sort(np.diag(Gm))[::-1])
    assert np.all(np.diff(np.sqrt(np.diag(Gm))) <= 1e-12)              # sorted descending as LAPACK returns them
    # truncate! with cutoff 1e-10 kept them all: the discarded tail is below the cutoff only if nothing smaller exists
    assert sig[-1] ** 2 / np.sum(sig ** 2) > 1e-10


def test_weighted_median_and_backwards_imputation(oracle):
    """Self-checks of the a18 restatements: StatsBase's weighted quantile on hand-computable cases, and
    impute_order=:backwards == the forward walk on the mirrored chain (an independent formulation: the oracle walks
    right to left after a left-to-right QR sweep; the mirrored chain uses the forward code)."""
    assert oracle.weighted_median([1.0, 2.0, 3.0], [1.0, 1.0, 1.0]) == 2.0
    assert abs(oracle.weighted_median([3.0, 1.0, 2.0], [0.2, 0.2, 0.6]) - 5.0 / 3.0) < 1e-15    # h = 0.6: 1 + (0.6-0.2)/0.6 * (2-1)
    assert abs(oracle.weighted_median([0.0, 1.0], [1.0, 1.0]) - 0.5) < 1e-15                     # h = 1.5: halfway 0 -> 1
    assert oracle.weighted_median([5.0, 0.0, 7.0], [0.0, 1.0, 0.0]) == 0.0                        # zero weights dropped
    N, T, d, C = 60, 9, 3, 2
    X, y = oracle.synthetic_two_class(N, T, seed=4)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.fit_sweeps(oracle.random_start_mps(T, d, 3, C, seed=2), phi, counts, nsweeps=2, chi_max=6, eta=0.05)
    cls = oracle.expand_label_index(cores)[0]
    mirrored = [A.transpose(2, 1, 0).copy() for A in reversed(cls)]
    grid = oracle.make_grid((-1.0, 1.0), 2e-3)
    genc = oracle.encode(grid, d)
    x = Xs[:, 5].copy()
    for ms in ([2, 3, 4], [0, 1], [6, 7, 8], [1, 4, 6]):
        for method in ("median", "mean", "mode"):
            a, _, ea, _ = oracle.impute_series_ex(cls, x, ms, grid, genc, d, method=method, impute_order="backwards", get_err=True)
            b, _, eb, _ = oracle.impute_series_ex(mirrored, x[::-1], [T - 1 - m for m in ms], grid, genc, d, method=method, get_err=True)
            assert np.abs(a - b[::-1]).max() < 1e-9 and np.abs(ea - eb[::-1]).max() < 1e-9, (ms, method)
    # forwards, default arguments: unchanged wrapper
    a, ia = oracle.impute_series(cls, x, [2, 3, 4], grid, genc, d)
    b, ib, _, _ = oracle.impute_series_ex(cls, x, [2, 3, 4], grid, genc, d)
    assert np.array_equal(a, b) and np.array_equal(ia, ib)


def test_imputation_with_per_site_encoders(oracle):
    """impute_at! with per-site encoders (time-dependent encodings: `xvals_enc[site]`, imputation.jl:92-100).  A per-site
    encoder that is the same built-in basis at every site must reproduce the built-in path exactly; a genuinely
    site-dependent one (the basis mirrored x -> -x on odd sites) must equal the built-in path on the model whose odd
    cores carry the same mirror (P_l(-x) = (-1)^l P_l(x))."""
    N, T, d, C = 80, 8, 4, 2
    X, y = oracle.synthetic_two_class(N, T, seed=9)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.fit_sweeps(oracle.random_start_mps(T, d, 4, C, seed=3), phi, counts, nsweeps=2, chi_max=6, eta=0.05)
    cls = oracle.expand_label_index(cores)[0]
    grid = oracle.make_grid((-1.0, 1.0), 2e-3)
    genc = oracle.encode(grid, d)
    x = Xs[:, order][:, 3]
    ms = [2, 3, 4, 6]
    U = np.array([0.3, 0.6, 0.2, 0.8])
    same = lambda j, v: oracle.encode(v, d)                                     # noqa: E731
    sign = (-1.0) ** np.arange(d)
    flip = lambda j, v: oracle.encode(v, d) * (sign if j % 2 else 1.0)          # noqa: E731  == encode(-v) on odd sites
    cls_flip = [A * (sign[None, :, None] if j % 2 else 1.0) for j, A in enumerate(cls)]
    genc_flip = np.stack([genc * (sign if j % 2 else 1.0) for j in range(T)])
    for method in ("median", "mean", "mode", "ITS"):
        for order_ in ("forwards", "backwards"):
            a = oracle.impute_series_ex(cls, x, ms, grid, genc, d, method=method, uniforms=U, impute_order=order_, get_err=True)
            b = oracle.impute_series_ex(cls, x, ms, grid, np.stack([genc] * T), d, basis=same, method=method, uniforms=U,
                                        impute_order=order_, get_err=True)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
            c = oracle.impute_series_ex(cls_flip, x, ms, grid, genc_flip, d, basis=flip, method=method, uniforms=U,
                                        impute_order=order_, get_err=True)
            assert np.abs(a[0] - c[0]).max() < 1e-9 and np.abs(a[2] - c[2]).max() < 1e-9
