"""Parity of the code paths the BENCHMARK actually runs (VERDICT r01 "the benchmarked code paths are not the tested
code paths"): bond shapes of BASELINE.json configs B (d=12, chi=40) and C (d=16, chi=64) from a trained-like state
(decaying spectra), the subspace-iteration SVD fast path with and without its column-scaling shortcut, a
cutoff-decided truncation inside that path, the register-operand K6 kernel for 17..48 outputs, K8 at the config-D
instance shape, and the reference's own trained ECG200 MPS.  Every test asserts WHICH kernel / path ran
(mpst_debug_get) so a silent fallback cannot make it pass.  Tolerances: BASELINE.json north_star."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SVD_SUBSPACE, GRAD_KR, GRAD_TILES, KRAO_REG, KRAO_TILES, KRAO_SLAB = 3, 1, 2, 1, 2, 3


def _device_trained(ctx, oracle, pkg, N, T, d, chi_max, nsweeps, eta=0.05, seed=1):
    """Train on the DEVICE (cheap), so the bonds compared afterwards have trained, decaying spectra."""
    X, y = oracle.synthetic_two_class(N, T, seed=seed)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.random_start_mps(T, d, 4, 2, seed=3)
    ctx.train_load_x(Xs[:, order], counts, d, chi_max)
    ctx.set_cores(cores)
    opts = pkg.make_opts(chi_max=chi_max, eta=eta)
    ctx.sweep_bonds(opts, nsweeps * 2 * (T - 1), restart=True, record=False)     # whole cycles: label back on site T-1
    return phi, counts, opts


def _oracle_bond(oracle, cs, phi, counts, j, going_left, chi_max, eta):
    N, T = phi.shape[0], phi.shape[1]
    L = np.ones((N, 1))
    for k in range(j):
        L = oracle.env_step_left(phi[:, k], L, cs[k])
    R = np.ones((N, 1))
    for k in range(T - 1, j + 1, -1):
        R = oracle.env_step_right(phi[:, k], R, cs[k])
    B, dims = oracle.flatten_bt(cs[j], cs[j + 1])
    Bn, lo, gn = oracle.apply_update(B, L, R, phi[:, j], phi[:, j + 1], counts, eta=eta)
    cl, cr, S = oracle.decompose_bt(Bn, dims, going_left, chi_max, 1e-10)
    return lo, gn, cl, cr, S


@pytest.mark.parametrize("N,T,d,chi,grad_kernel,bonds", [
    (2048, 10, 12, 40, GRAD_KR, (3, 4, 5, 6)),        # config B bond shape: register-operand gradient, krao_slab<NI=5>, p = 80
    (1024, 8, 16, 64, None, (2, 3, 4)),               # config C (north star) bond shape: 2048 x 1024 split, p = 112
])
def test_teacher_forced_bonds_at_benchmark_shapes(ctx, oracle, pkg, N, T, d, chi, grad_kernel, bonds):
    """The device walks a backward half-sweep from a device-trained state; before each compared bond its cores are
    downloaded and the ORACLE performs the same bond from the same state (teacher forcing by construction)."""
    phi, counts, opts = _device_trained(ctx, oracle, pkg, N, T, d, chi, nsweeps=2)
    eta = 0.05
    seen_subspace = 0
    for j in range(T - 2, min(bonds) - 1, -1):
        if j not in bonds:
            ctx.bond_step(j, True, opts)
            continue
        cs = ctx.get_cores()
        assert cs[j].shape[0] == chi and cs[j + 1].shape[2] == chi, "bond is not chi-saturated: pick other bonds"
        lo, gn, k = ctx.bond_step(j, True, opts)
        path, gk, kk, kv = (ctx.debug_get(n) for n in ("svd_path", "grad_kernel", "krao_kernel", "krao_variant"))
        lo_r, gn_r, cl, cr, S = _oracle_bond(oracle, cs, phi, counts, j, True, chi, eta)
        assert abs(lo - lo_r) <= RTOL * abs(lo_r), (j, lo, lo_r)
        assert abs(gn - gn_r) <= RTOL * gn_r, (j, gn, gn_r)
        assert k == len(S), (j, k, len(S))
        dl, dr = ctx.get_core(j), ctx.get_core(j + 1)
        prod_d = np.einsum("asmc,mtb->btasc", dl, dr)
        prod_r = np.einsum("asmc,mtb->btasc", cl, cr)
        assert np.abs(prod_d - prod_r).max() < 1e-10, (j, np.abs(prod_d - prod_r).max())
        sig_d = np.linalg.svd(dl.transpose(0, 3, 1, 2).reshape(-1, k), compute_uv=False)    # label core = U S
        assert np.abs(sig_d - S).max() < 1e-10 * S.max()
        orth = dr.reshape(k, -1)
        assert np.abs(orth @ orth.T - np.eye(k)).max() < 1e-10
        # the paths the benchmark takes at this shape, not their fallbacks
        assert path == SVD_SUBSPACE, f"bond {j}: SVD took path {path}, not the subspace iteration"
        seen_subspace += 1
        if grad_kernel is not None:
            assert gk == grad_kernel
        assert kk == KRAO_SLAB and kv // 100 == (chi + 7) // 8      # krao_slab_kernel, chi/8 column fragments
    assert seen_subspace == len(bonds)


def test_column_scaling_shortcut_equals_full_orthonormalisation(ctx, oracle, pkg):
    """svd_subspace.cu's shortcut (Z = M Q only column-normalised after the first iteration) must not change the split:
    same chi and the same truncated two-site product to 1e-12 on every saturated bond of a half-sweep.  Each bond is
    teacher-forced from the state the reference walk had before it (free-running KLD amplifies rounding)."""
    N, T, d, chi = 1024, 10, 12, 40
    phi, counts, opts = _device_trained(ctx, oracle, pkg, N, T, d, chi, nsweeps=1)
    states, prods, chis, shortcut_bonds = {}, {}, {}, 0
    for j in range(T - 2, 1, -1):                            # reference walk: shortcut on (the default)
        states[j] = ctx.get_cores()
        _, _, chis[j] = ctx.bond_step(j, True, opts)
        if states[j][j].shape[0] == chi and states[j][j + 1].shape[2] == chi:
            assert ctx.debug_get("svd_path") == SVD_SUBSPACE and ctx.debug_get("svd_restarts") == 0
            shortcut_bonds += 1
        prods[j] = np.einsum("asmc,mtb->btasc", ctx.get_core(j), ctx.get_core(j + 1))
    assert shortcut_bonds >= 4
    ctx.debug_set("SVD_NOHALF", 1)
    try:
        for j in range(T - 2, 1, -1):
            ctx.set_cores(states[j])
            ctx.build_env(True)
            ctx.build_env(False)
            _, _, k = ctx.bond_step(j, True, opts)
            assert k == chis[j]
            p = np.einsum("asmc,mtb->btasc", ctx.get_core(j), ctx.get_core(j + 1))
            assert np.abs(p - prods[j]).max() < 1e-12, (j, np.abs(p - prods[j]).max())
    finally:
        ctx.debug_set("SVD_NOHALF", 0)


def _decaying_bond(rng, d, cl, cr, C, knee, r1, r2):
    """Full-rank bond matrix with the two-slope spectrum of a trained bond: s_k = r1^k up to the knee, then the slow
    tail r1^knee * r2^(k - knee) (what the gradient term adds).  Unit Frobenius norm, returned as a (D, C) bond tensor
    (going left: rows (a, c, s_l), columns (s_r, b))."""
    m, n = cl * C * d, d * cr
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    k = np.arange(n)
    s = np.where(k < knee, r1 ** np.minimum(k, knee), r1 ** knee * r2 ** np.maximum(k - knee, 0))
    M = (U * s) @ V.T
    M /= np.linalg.norm(M)
    full = M.reshape(cl, C, d, d, cr)                      # (a, c, s_l, s_r, b)
    B = full.transpose(1, 4, 3, 0, 2).reshape(C, -1).T     # column c: index s_l + d*(a + cl*(s_r + d*b))
    return np.ascontiguousarray(B)


@pytest.mark.parametrize("d,chi,C,knee,r1,r2,chimax,cutoff_decides", [
    (12, 40, 2, 20, 0.5, 0.99, 40, True),      # 960 x 480: the 1e-10 cutoff stops at ~18 < chi_max
    (12, 40, 2, 26, 0.6, 0.99, 40, True),      # slower head: ~25 kept, convergence factor 0.04 per iteration
    (16, 64, 2, 20, 0.5, 0.995, 64, True),     # 2048 x 1024 (north-star split), cutoff-decided
    (12, 40, 2, 0, 0.9, 0.9, 40, False),       # slow geometric decay: chi_max decides
    (16, 64, 2, 0, 0.93, 0.93, 64, False),
])
def test_subspace_svd_cutoff_decided_truncation(ctx, oracle, d, chi, C, knee, r1, r2, chimax, cutoff_decides):
    """The truncation rule inside the subspace path: kept dimension, singular values and product equal LAPACK +
    NDTensors' truncate! as restated by the oracle, and the fast path (not the Jacobi fallback) produced them."""
    rng = np.random.default_rng(int(1000 * r1) + d)
    B = _decaying_bond(rng, d, chi, chi, C, knee, r1, r2)
    c_l, c_r, sig = ctx.bond_split(B, d, chi, chi, True, chimax)
    path = ctx.debug_get("svd_path")
    r_l, r_r, rs = oracle.decompose_bt(B, (chi, d, chi), True, chimax, 1e-10)
    assert len(sig) == len(rs), (len(sig), len(rs))
    assert (len(rs) < chimax) == cutoff_decides, len(rs)
    assert np.abs(sig - rs).max() < 1e-10 * rs.max()
    ein = "asmc,mtb->btasc"
    assert np.abs(np.einsum(ein, c_l, c_r) - np.einsum(ein, r_l, r_r)).max() < 1e-10
    assert path == SVD_SUBSPACE, f"took path {path}"
    # the cutoff is a RELATIVE tail weight: the same matrix scaled by 7 splits identically (sigma scales, chi does not)
    _, _, sig7 = ctx.bond_split(7.0 * B, d, chi, chi, True, chimax)
    assert len(sig7) == len(rs) and np.abs(sig7 - 7.0 * rs).max() < 1e-9 * rs.max()


@pytest.mark.parametrize("d,chi,C,rank", [(12, 40, 2, 24), (16, 64, 2, 40), (12, 40, 2, 70)])
def test_subspace_svd_rank_deficient_bond(ctx, oracle, d, chi, C, rank):
    """A numerically rank-deficient bond (a chi_init start, or a tail below rounding level): the Cholesky-QR steps deflate
    the dependent columns instead of abandoning the subspace path; kept dimension / sigma / product as LAPACK."""
    rng = np.random.default_rng(rank)
    m, n = chi * C * d, d * chi
    U, _ = np.linalg.qr(rng.standard_normal((m, rank)))
    V, _ = np.linalg.qr(rng.standard_normal((n, rank)))
    M = (U * (0.8 ** np.arange(rank))) @ V.T
    M /= np.linalg.norm(M)
    B = np.ascontiguousarray(M.reshape(chi, C, d, d, chi).transpose(1, 4, 3, 0, 2).reshape(C, -1).T)
    c_l, c_r, sig = ctx.bond_split(B, d, chi, chi, True, chi)
    path, restarts = ctx.debug_get("svd_path"), ctx.debug_get("svd_restarts")
    r_l, r_r, rs = oracle.decompose_bt(B, (chi, d, chi), True, chi, 1e-10)
    assert len(sig) == len(rs) <= min(rank, chi), (len(sig), len(rs))
    assert np.abs(sig - rs).max() < 1e-10 * rs.max()
    ein = "asmc,mtb->btasc"
    assert np.abs(np.einsum(ein, c_l, c_r) - np.einsum(ein, r_l, r_r)).max() < 1e-10
    assert path == SVD_SUBSPACE and restarts == 0, (path, restarts)


def test_svd_graph_replay_equals_plain_launches(ctx, oracle):
    """The captured CUDA graph of a subspace round replays exactly the launches it recorded: bit-identical cores with
    the graph, without it, and with the serial (single-stream) loop's Gram / Cholesky order where that applies."""
    rng = np.random.default_rng(5)
    d, chi, C = 12, 40, 2
    B = _decaying_bond(rng, d, chi, chi, C, 0, 0.9, 0.9)
    outs = []
    try:
        for nograph in (0, 1, 0):
            ctx.debug_set("SVD_NOGRAPH", nograph)
            outs.append(ctx.bond_split(B, d, chi, chi, True, chi))
            assert ctx.debug_get("svd_path") == SVD_SUBSPACE
    finally:
        ctx.debug_set("SVD_NOGRAPH", 0)
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(outs[0], o))


@pytest.mark.parametrize("d,chi", [(16, 64), (12, 64), (8, 52), (16, 128), (12, 40), (24, 32)])
def test_krao_slab_kernel(ctx, oracle, d, chi):
    """K6/K7 for wide outputs: krao_slab_kernel (W streamed in K-slabs, register operands) vs the oracle and vs the
    shared-memory tile kernel it replaces (MPST_KRAO_NOSLAB), both row-block sizes."""
    N, T, C = 1500, 4, 2
    rng = np.random.default_rng(d * chi)
    X = rng.uniform(-1, 1, (T, N))
    phi = oracle.encode(X.T, d)
    cores = oracle.random_start_mps(T, d, chi, C, seed=chi)
    ctx.model_init(T, C, d, chi)
    ctx.set_cores(cores)
    ref = oracle.overlaps(cores, phi)
    outs = {}
    for name, flags in (("slab4", {"KRAO_NOSLAB": 0, "KRAO_SLAB_MI": 4}), ("slab2", {"KRAO_NOSLAB": 0, "KRAO_SLAB_MI": 2}),
                        ("tiles", {"KRAO_NOSLAB": 1, "KRAO_SLAB_MI": 0})):
        for k, v in flags.items():
            ctx.debug_set(k, v)
        ctx.debug_set("krao_slab_launches", 0)
        yh, am = ctx.overlaps(X_TxN=X)
        outs[name] = (yh, ctx.debug_get("krao_slab_launches"))
        assert np.abs(yh - ref).max() < 1e-12 * np.abs(ref).max(), name
        assert np.array_equal(am, np.argmax(ref * ref, axis=1)), name
    ctx.debug_set("KRAO_NOSLAB", 0)
    ctx.debug_set("KRAO_SLAB_MI", 0)
    assert outs["slab4"][1] > 0 and outs["slab2"][1] > 0 and outs["tiles"][1] == 0
    assert np.abs(outs["slab4"][0] - outs["tiles"][0]).max() < 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("chi", [17, 24, 32, 40, 48])
def test_krao_reg_kernel_all_widths(ctx, oracle, chi):
    """K6/K7 register-operand kernel krao_reg_kernel<NI> for NI = 3..6 (17..48 outputs): environments through a
    uniform-chi chain vs the oracle."""
    ctx.debug_set("KRAO_SLAB_MIN", 1000)          # the slab kernel would take chi = 32, 40, 48
    try:
        N, T, d, C = 600, 5, 8, 2
        rng = np.random.default_rng(chi)
        X = rng.uniform(-1, 1, (T, N))
        phi = oracle.encode(X.T, d)
        cores = oracle.random_start_mps(T, d, chi, C, seed=chi)
        ctx.model_init(T, C, d, chi)
        ctx.set_cores(cores)
        yh, am = ctx.overlaps(X_TxN=X)
        ref = oracle.overlaps(cores, phi)
        assert np.abs(yh - ref).max() < 1e-12 * np.abs(ref).max()
        assert np.array_equal(am, np.argmax(ref * ref, axis=1))
        # the training-side environment chain of the same model (mpst_build_env -> K6) takes the register kernel
        counts = np.array([N // 2, N - N // 2])
        ctx.train_load_x(X, counts, d, chi)
        ctx.set_cores(cores)
        ctx.debug_set("krao_reg_mask", 0)
        ctx.build_env(True)
        assert ctx.debug_get("krao_reg_mask") & (1 << ((chi + 7) // 8)), ctx.debug_get("krao_reg_mask")
        lo, gn, k = ctx.bond_step(T - 2, True, ctx_opts(chi))
        L = np.ones((N, 1))
        for j in range(T - 2):
            L = oracle.env_step_left(phi[:, j], L, cores[j])
        B, dims = oracle.flatten_bt(cores[T - 2], cores[T - 1])
        lo_r, G_r = oracle.loss_grad_KLD(B, L, np.ones((N, 1)), phi[:, T - 2], phi[:, T - 1], counts)
        assert abs(lo - lo_r) <= RTOL * abs(lo_r) and abs(gn - np.linalg.norm(G_r)) <= RTOL * np.linalg.norm(G_r)
    finally:
        ctx.debug_set("KRAO_SLAB_MIN", 25)


def ctx_opts(chi):
    import mpstime_jl_b200 as m
    return m.make_opts(chi_max=chi, eta=0.05)


def test_reference_trained_mps_golden(ctx, pkg, oracle):
    """The reference's OWN trained MPS (test/Data/ecg200/mps_saves/test_dataset.jld2, extracted by
    tests/golden/make_golden_mps_from_jld2.py): K7 on the device reproduces <W|W> = 1 (normalize!, :852), the label on
    the last site, and classifies the reference's own encoded training set like the CPU contraction of the same cores
    (100 % -- the fixture was trained to convergence on these 100 series)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ecg200_trained_mps.npz"))
    e = np.load(os.path.join(ROOT, "tests", "golden", "ecg200_legendre.npz"))
    T = 96
    cores = [g["core_%02d" % j] for j in range(T)]
    assert int(g["label_pos"]) == T - 1 and cores[-1].ndim == 4
    chi = max(max(A.shape[0], A.shape[2]) for A in cores)
    y = np.r_[np.zeros(31, dtype=np.int64), np.ones(69, dtype=np.int64)]          # class_distribution of the fixture
    # (1) precomputed-phi entry on the reference's own PStates
    ctx.train_load_phi(e["phi_ref"], np.array([31, 69]), chi)
    ctx.set_cores(cores)
    yh, am = ctx.overlaps(phi_NTd=e["phi_ref"])
    ref = oracle.overlaps(cores, e["phi_ref"])
    assert np.abs(yh - ref).max() < 1e-12 * np.abs(ref).max()
    assert np.array_equal(am, y)
    # (2) the public classify path from RAW series: normalisation + on-device Legendre encoding + contraction
    opts = pkg.MPSOptions(d=5, chi_max=chi)
    order_data = pkg.EncodedTimeSeriesSet(None, e["X_orig"], y, np.array([31, 69]))
    mps = pkg.TrainedMPS(cores, opts, order_data, classes=np.array([0, 1]))
    assert np.array_equal(pkg.classify(mps, e["X_orig"]), y)
    # (3) the stored state is what our sweep would leave: unit norm and left-canonical up to the label site
    back = ctx.get_cores()
    assert all(np.array_equal(a, b) for a, b in zip(cores, back))
    assert abs(oracle._norm2_general(cores) - 1.0) < 1e-12


def test_k8_config_d_instance_shape(ctx, oracle, pkg):
    """One MPS_impute instance at BASELINE configs[3] shape: T = 256, d = 16, chi_max = 64, 128 contiguous missing
    sites, the dx = 1e-4 grid (G = 20 001), median and ITS, against the oracle within 1e-8.  The class MPS is trained
    on the device (one sweep on 512 series) so the conditionals are those of a trained model."""
    T, d, chi, K = 256, 16, 64, 128
    rng = np.random.default_rng(11)
    t = np.arange(1, T + 1)
    Xtr = np.sin(2 * np.pi * t[None, :] / 24.0 + rng.uniform(0, 2 * np.pi, 512)[:, None]) + 0.2 * rng.standard_normal((512, T))
    opts = pkg.MPSOptions(d=d, chi_max=chi, nsweeps=1, eta=0.05, verbosity=-1, log_level=0)
    mps, _, _ = pkg.fitMPS(Xtr, np.zeros(512, dtype=np.int64), opts=opts)
    cores = mps.mps
    assert max(A.shape[2] for A in cores) == chi
    Xs, norms = oracle.transform_train_data(Xtr.T)
    grid = oracle.make_grid((-1.0, 1.0), 1e-4)
    assert len(grid) == 20001
    genc = oracle.encode(grid, d)
    ctx.model_init(T, 1, d, chi)
    ctx.set_cores(cores)
    n = 3
    X = Xs[:, :n].copy()
    mask = np.zeros((T, n), dtype=np.uint8)
    starts = [0, 64, 128]                                    # left edge, bulk, right edge
    for k, s0 in enumerate(starts):
        mask[s0:s0 + K, k] = 1
        X[s0:s0 + K, k] = 0.0
    U = rng.uniform(0.05, 0.95, size=(n, 1, K))
    cls = oracle.expand_label_index(cores)[0]
    for method in ("median", "ITS"):
        out = ctx.impute_batch(0, X, mask, grid, method=method, uniforms=U if method == "ITS" else None)
        for k, s0 in enumerate(starts):
            ms = list(range(s0, s0 + K))
            ref, _ = oracle.impute_series(cls, X[:, k], ms, grid, genc, d, method=method,
                                          uniforms=U[k, 0] if method == "ITS" else None)
            assert np.abs(out[k, 0, ms] - ref[ms]).max() < 1e-8, (method, k, np.abs(out[k, 0, ms] - ref[ms]).max())


def test_non_finite_loss_is_reported_without_outputs(ctx, oracle, pkg):
    """a zero overlap (KLD weight -1/(N yhat)) must surface as MPST_E_NUMERIC on the quiet path too."""
    N, T, d = 130, 5, 3
    X, y = oracle.synthetic_two_class(N, T, seed=2)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.random_start_mps(T, d, 4, 2, seed=3)
    cores[1] = np.zeros_like(cores[1])                       # every overlap is exactly zero
    ctx.train_load_x(Xs[:, order], counts, d, 6)
    ctx.set_cores(cores)
    ctx.build_env(True)
    with pytest.raises(pkg.MPSTError, match="non-finite"):
        ctx.bond_step_quiet(T - 2, True, pkg.make_opts(chi_max=6, eta=0.05))


def test_stale_or_mirrored_environment_is_refused(ctx, oracle, pkg):
    """bond_step must not contract a stale or wrong-direction environment when the link dimensions match."""
    N, T, d = 130, 6, 3
    X, y = oracle.synthetic_two_class(N, T, seed=2)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.random_start_mps(T, d, 3, 2, seed=3)
    opts = pkg.make_opts(chi_max=3, eta=0.05)
    ctx.train_load_x(Xs[:, order], counts, d, 3)
    ctx.set_cores(cores)
    ctx.build_env(True)
    ctx.set_core(1, 1.5 * cores[1])                          # LE[1..] are now stale although every dimension matches
    with pytest.raises(pkg.MPSTError, match="stale"):
        ctx.bond_step(T - 2, True, opts)
    ctx.build_env(True)
    ctx.bond_step(T - 2, True, opts)                         # fine again


def test_sweep_bonds_equals_sweep(ctx, oracle, pkg):
    """mpst_sweep_bonds is the sweep loop cut into pieces: same per-bond records, bit-identical cores before normalize!."""
    N, T, d = 300, 7, 4
    X, y = oracle.synthetic_two_class(N, T, seed=4)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.random_start_mps(T, d, 4, 2, seed=3)
    opts = pkg.make_opts(chi_max=9, eta=0.05)
    ctx.train_load_x(Xs[:, order], counts, d, 9)
    ctx.set_cores(cores)
    lo, gn, chi = ctx.sweep(opts, 2)
    ref = ctx.get_cores()
    ctx.train_load_x(Xs[:, order], counts, d, 9)
    ctx.set_cores(cores)
    nb = 2 * 2 * (T - 1)
    parts = [ctx.sweep_bonds(opts, 5, restart=True)]
    done = 5
    while done < nb:
        k = min(7, nb - done)
        parts.append(ctx.sweep_bonds(opts, k))
        done += k
    lo2, gn2, chi2 = (np.concatenate([p[i] for p in parts]) for i in range(3))
    assert np.array_equal(lo, lo2) and np.array_equal(gn, gn2) and np.array_equal(chi, chi2)
    mine = pkg.api._normalize(ctx.get_cores())
    for a, b in zip(ref, mine):
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-12 * max(1.0, np.abs(a).max())
