"""Parity of the CUDA path against the CPU oracle, through the C ABI (ctypes).  Run with -m gpu.
Tolerances follow BASELINE.json north_star: relative 1e-9 on per-bond loss and gradient norm,
identical truncated bond dimensions, identical class predictions."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rand_bond(oracle, rng, N, d, chi_l, chi_r, C, counts=None):
    if counts is None:
        counts = [N // C] * C
        counts[-1] += N - sum(counts)
    xl = oracle.legendre_encode(rng.uniform(-1, 1, N), d)
    xr = oracle.legendre_encode(rng.uniform(-1, 1, N), d)
    L = rng.standard_normal((N, chi_l)) / np.sqrt(chi_l) if chi_l > 1 else np.ones((N, 1))
    R = rng.standard_normal((N, chi_r)) / np.sqrt(chi_r) if chi_r > 1 else np.ones((N, 1))
    B = rng.standard_normal((d * chi_l * d * chi_r, C))
    B /= np.linalg.norm(B)
    return B, L, R, xl, xr, np.array(counts)


# ---- K1 -------------------------------------------------------------------------------------
@pytest.mark.parametrize("basis,d,unit", [("legendre_no_norm", 5, False), ("legendre_no_norm", 16, False),
                                          ("legendre_norm", 7, False), ("fourier", 6, False), ("fourier", 9, False),
                                          ("stoudenmire", 2, True), ("sahand", 6, True), ("uniform", 3, True)])
def test_encode_matches_oracle(ctx, oracle, basis, d, unit):
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, 4099) if unit else rng.uniform(-1, 1, 4099)
    x[:3] = [0.0, 1.0, 0.5] if unit else [-1.0, 1.0, 0.0]
    dev = ctx.encode(x, d, basis)
    ref = oracle.encode(x, d, basis)
    assert dev.shape == ref.shape and dev.dtype == ref.dtype
    assert np.abs(dev - ref).max() < 1e-13


def test_encode_golden_reference_output(ctx, pkg):
    """Device Legendre encoding vs the reference's own saved PStates (ECG200, d=5)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ecg200_legendre.npz"))
    Xs, _ = pkg.transform_train_data(g["X_orig"].T, pkg.MPSOptions())
    dev = ctx.encode(Xs.T, 5, "legendre_no_norm")
    assert np.abs(dev - g["phi_ref"]).max() < 1e-14


def test_encode_empty_and_bad_args(ctx, pkg):
    assert ctx.encode(np.zeros(0), 4).shape == (0, 4)
    with pytest.raises(pkg.MPSTError):
        ctx.encode(np.zeros(3), 3, "stoudenmire")
    with pytest.raises(pkg.MPSTError):
        ctx.encode(np.zeros(3), 5, "sahand")


# ---- K2: loss + gradient, teacher-forced on identical operands --------------------------------
SHAPES = [  # N, d, chi_l, chi_r, C, counts
    (300, 4, 3, 5, 2, None),
    (1000, 12, 40, 40, 2, None),          # BASELINE config B bond shape
    (257, 5, 1, 7, 3, [100, 1, 156]),     # first-site bond (no LE), ragged classes, one singleton class
    (100, 6, 9, 1, 1, None),              # last-site bond (no RE), single class (unsupervised)
    (64, 2, 1, 1, 2, None),               # two-site MPS
    (1500, 16, 64, 64, 2, [1499, 1]),     # north-star bond shape, extreme imbalance
    (513, 10, 20, 20, 2, None),           # BASELINE config A shape
    (700, 24, 16, 16, 2, None),           # micro-sweep corner d=24
    (400, 12, 10, 14, 2, None),           # register-operand kernel, partial 8-wide link blocks on both sides
    (333, 6, 8, 22, 3, [300, 2, 31]),     # one site group per side, ragged classes, chunk-straddling class edges
    (200, 8, 12, 30, 2, None),            # d % 4 variant (2 link blocks x 4 sites per warp)
    (50, 18, 8, 8, 2, [49, 1]),           # fewer samples than one stream-K range per SM
]


@pytest.mark.parametrize("N,d,cl,cr,C,counts", SHAPES)
@pytest.mark.parametrize("loss,sep", [("KLD", False), ("KLD", True), ("MSE", False)])
def test_bond_loss_grad(ctx, oracle, N, d, cl, cr, C, counts, loss, sep):
    rng = np.random.default_rng(N + d)
    B, L, R, xl, xr, counts = rand_bond(oracle, rng, N, d, cl, cr, C, counts)
    lo, G, yh = ctx.bond_loss_grad(B, L, R, xl, xr, counts, loss=loss, train_sep=sep, want_yhat=True)
    if loss == "KLD":
        lo_r, G_r = oracle.loss_grad_KLD(B, L, R, xl, xr, counts, sep)
    else:
        lo_r, G_r = oracle.loss_grad_MSE(B, L, R, xl, xr, counts)
    assert abs(lo - lo_r) <= RTOL * abs(lo_r)
    assert abs(np.linalg.norm(G) - np.linalg.norm(G_r)) <= RTOL * np.linalg.norm(G_r)
    assert np.abs(G - G_r).max() <= 1e-9 * np.abs(G_r).max()
    yr = oracle.bond_yhat(B, L, R, xl, xr)
    if loss == "KLD":
        own = np.repeat(np.arange(C), counts)
        assert np.abs(yh[np.arange(N), own] - yr[np.arange(N), own]).max() < 1e-12 * np.abs(yr).max() + 1e-15
    else:
        assert np.abs(yh - yr).max() < 1e-12 * np.abs(yr).max() + 1e-15


def test_bond_loss_grad_vs_literal_reference_loop(ctx, oracle):
    """against the sample-sequential, phi~-explicit form the reference actually executes."""
    rng = np.random.default_rng(7)
    B, L, R, xl, xr, counts = rand_bond(oracle, rng, 90, 3, 4, 5, 2, [40, 50])
    lo, G = ctx.bond_loss_grad(B, L, R, xl, xr, counts)
    lo_r, G_r = oracle.loss_grad_KLD_loop(B, L, R, xl, xr, counts)
    assert abs(lo - lo_r) <= RTOL * abs(lo_r) and np.abs(G - G_r).max() <= 1e-9 * np.abs(G_r).max()


def test_gradient_linearity_at_scale(ctx, oracle):
    """size-independent property at a BASELINE-sized bond: the MSE gradient is affine in B and the
    KLD gradient is homogeneous of degree -1: G(a*B) = G(B)/a, loss(a*B) = loss(B) - log(a^2)."""
    rng = np.random.default_rng(3)
    N, d, chi = 20000, 12, 40
    B, L, R, xl, xr, counts = rand_bond(oracle, rng, N, d, chi, chi, 2)
    lo1, G1 = ctx.bond_loss_grad(B, L, R, xl, xr, counts)
    lo2, G2 = ctx.bond_loss_grad(3.0 * B, L, R, xl, xr, counts)
    assert abs(lo2 - (lo1 - np.log(9.0))) < 1e-10 * abs(lo1)
    assert np.abs(3.0 * G2 - G1).max() < 1e-10 * np.abs(G1).max()
    # permuting the samples inside a class does not change the result beyond rounding
    perm = np.concatenate([rng.permutation(counts[0]), counts[0] + rng.permutation(counts[1])])
    lo3, G3 = ctx.bond_loss_grad(B, L[perm], R[perm], xl[perm], xr[perm], counts)
    assert abs(lo3 - lo1) < 1e-11 * abs(lo1) and np.abs(G3 - G1).max() < 1e-10 * np.abs(G1).max()


# ---- K5: truncated SVD split -----------------------------------------------------------------
@pytest.mark.parametrize("d,cl,cr,C,gl,chimax", [(4, 3, 5, 2, True, 8), (4, 3, 5, 2, False, 8), (12, 40, 40, 2, True, 40),
                                                 (12, 40, 40, 2, False, 40), (5, 1, 6, 2, True, 10), (5, 6, 1, 2, False, 10),
                                                 (3, 2, 2, 3, True, 100), (16, 64, 64, 2, False, 64), (10, 20, 1, 2, True, 20)])
def test_bond_split_matches_lapack(ctx, oracle, d, cl, cr, C, gl, chimax):
    rng = np.random.default_rng(d * cl + cr)
    B = rng.standard_normal((d * cl * d * cr, C))
    B /= np.linalg.norm(B)
    c_l, c_r, sig = ctx.bond_split(B, d, cl, cr, gl, chimax)
    r_l, r_r, rs = oracle.decompose_bt(B, (cl, d, cr), gl, chimax, 1e-10)
    assert len(sig) == len(rs)                                   # identical truncated bond dimension
    assert np.abs(sig - rs).max() < 1e-10 * rs.max()
    ein = "asmc,mtb->btasc" if gl else "asm,mtbc->btasc"
    assert np.abs(np.einsum(ein, c_l, c_r) - np.einsum(ein, r_l, r_r)).max() < 1e-10    # gauge-invariant product
    orth = c_r.reshape(c_r.shape[0], -1) if gl else c_l.reshape(-1, c_l.shape[2]).T
    assert np.abs(orth @ orth.T - np.eye(orth.shape[0])).max() < 1e-10


def test_bond_split_cutoff_truncation(ctx, oracle):
    """low-rank bond + tiny noise: the relative cutoff (not chi_max) decides; same decision as the oracle."""
    rng = np.random.default_rng(5)
    d, cl, cr, C = 4, 6, 6, 2
    a = rng.standard_normal((cl, d, 3, C))
    b = rng.standard_normal((3, d, cr))
    B = np.einsum("asmc,mtb->btasc", a, b).reshape(-1, C)
    B /= np.linalg.norm(B)
    B = B + 1e-9 * rng.standard_normal(B.shape)
    for cutoff in (1e-10, 1e-20, 1e-4):
        c_l, c_r, sig = ctx.bond_split(B, d, cl, cr, True, 20, cutoff)
        r_l, r_r, rs = oracle.decompose_bt(B, (cl, d, cr), True, 20, cutoff)
        assert len(sig) == len(rs), (cutoff, len(sig), len(rs))
    assert len(ctx.bond_split(B, d, cl, cr, True, 20, 1e-10)[2]) == 3


# ---- full path --------------------------------------------------------------------------------
def _problem(oracle, N, T, d, C=2, seed=1, chi_init=4):
    X, y = oracle.synthetic_two_class(N, T, seed=seed)
    Xs, _ = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.random_start_mps(T, d, chi_init, C, seed=3)
    return Xs[:, order], phi, ys, counts, cores


def test_core_roundtrip_and_env(ctx, oracle):
    Xs, phi, ys, counts, cores = _problem(oracle, 130, 7, 3)
    ctx.train_load_x(Xs, counts, 3, 8)
    ctx.set_cores(cores)
    back = ctx.get_cores()
    assert all(np.array_equal(a, b) for a, b in zip(cores, back))
    yh, am = ctx.overlaps(X_TxN=Xs)
    ref = oracle.overlaps(cores, phi)
    assert np.abs(yh - ref).max() < 1e-13 * np.abs(ref).max() + 1e-16
    assert np.array_equal(am, np.argmax(ref * ref, axis=1))


@pytest.mark.parametrize("kw", [dict(), dict(loss="MSE"), dict(bbopt="GD", eta=0.02), dict(update_iters=2),
                                dict(train_sep=True), dict(rescale=(True, False))])
def test_teacher_forced_bond_steps_through_the_sweep_path(ctx, oracle, pkg, kw):
    """Every bond of two sweeps: load the ORACLE's cores before the bond, rebuild the environments on
    the device, run mpst_bond_step, compare loss / gradient norm / chi / the truncated two-site product
    and the new environment (gauge-invariant)."""
    N, T, d, C, chi_max = 160, 6, 3, 2, 7
    Xs, phi, ys, counts, cores = _problem(oracle, N, T, d, C)
    okw = dict(chi_max=chi_max, eta=0.05)
    okw.update(kw)
    mkw = dict(okw)
    opts = pkg.make_opts(**mkw)
    cs = [c.copy() for c in cores]
    ones = np.ones((N, 1))
    for sweep in range(2):
        for going_left in (True, False):
            for j in (range(T - 2, -1, -1) if going_left else range(T - 1)):
                ctx.train_load_x(Xs, counts, d, chi_max)
                ctx.set_cores(cs)
                ctx.build_env(True)
                ctx.build_env(False)
                lo, gn, chi = ctx.bond_step(j, going_left, opts)
                # oracle step from the same state
                L = ones
                for k in range(j):
                    L = oracle.env_step_left(phi[:, k], L, cs[k])
                R = ones
                for k in range(T - 1, j + 1, -1):
                    R = oracle.env_step_right(phi[:, k], R, cs[k])
                B, dims = oracle.flatten_bt(cs[j], cs[j + 1])
                Bn, lo_r, gn_r = oracle.apply_update(B, L, R, phi[:, j], phi[:, j + 1], counts, loss=okw.get("loss", "KLD"),
                                                     bbopt=okw.get("bbopt", "TSGO"), eta=okw["eta"],
                                                     update_iters=okw.get("update_iters", 1),
                                                     rescale=okw.get("rescale", (False, True)),
                                                     train_sep=okw.get("train_sep", False))
                cl, cr, S = oracle.decompose_bt(Bn, dims, going_left, chi_max, 1e-10)
                assert abs(lo - lo_r) <= RTOL * abs(lo_r), (sweep, going_left, j)
                assert abs(gn - gn_r) <= RTOL * gn_r, (sweep, going_left, j)
                assert chi == len(S), (sweep, going_left, j)
                dl, dr = ctx.get_core(j), ctx.get_core(j + 1)
                ein = "asmc,mtb->btasc" if going_left else "asm,mtbc->btasc"
                assert np.abs(np.einsum(ein, dl, dr) - np.einsum(ein, cl, cr)).max() < 1e-10
                cs[j], cs[j + 1] = cl, cr


def test_free_running_sweep_matches_oracle(ctx, oracle, pkg):
    """Full sweeps without teacher forcing.  The KLD recursion is chaotic from a random start (the gradient
    is dominated by the samples whose overlap is closest to zero, weight 1/yhat), so rounding differences
    grow ~10x per bond (reference docs/src/classification.md:57-60): KLD is compared on the first
    bonds only; the well-conditioned MSE/GD recursion is compared over two whole sweeps."""
    N, T, d, C = 400, 8, 4, 2
    Xs, phi, ys, counts, cores = _problem(oracle, N, T, d, C)
    rec = []
    oracle.fit_sweeps(cores, phi, counts, nsweeps=1, chi_max=10, eta=0.05, record=rec)
    ctx.train_load_x(Xs, counts, d, 10)
    ctx.set_cores(cores)
    lo, gn, chi = ctx.sweep(pkg.make_opts(chi_max=10, eta=0.05), 1)
    rl = np.array([r["loss"] for r in rec])
    assert np.array_equal(chi, [r["chi"] for r in rec])
    assert np.abs(lo[:4] - rl[:4]).max() < 1e-9 * np.abs(rl).max()
    dev = ctx.get_cores()
    assert abs(oracle._norm2_general(dev) - 1.0) < 1e-10             # normalize!(W)
    yh, am = ctx.overlaps(X_TxN=Xs)
    yd = oracle.overlaps(dev, phi)
    assert np.abs(yh - yd).max() < 1e-12 * np.abs(yd).max()
    # MSE + GD: smooth recursion, whole trajectory and final predictions agree
    rec = []
    new = oracle.fit_sweeps(cores, phi, counts, nsweeps=2, chi_max=10, eta=0.3, loss="MSE", bbopt="GD", record=rec)
    ctx.train_load_x(Xs, counts, d, 10)
    ctx.set_cores(cores)
    lo, gn, chi = ctx.sweep(pkg.make_opts(chi_max=10, eta=0.3, loss="MSE", bbopt="GD"), 2)
    rl = np.array([r["loss"] for r in rec])
    assert np.array_equal(chi, [r["chi"] for r in rec])
    assert np.abs(lo - rl).max() < 1e-8 * np.abs(rl).max()
    dev = ctx.get_cores()
    yd, yr = oracle.overlaps(dev, phi), oracle.overlaps(new, phi)
    assert np.abs(np.abs(yd) - np.abs(yr)).max() < 1e-7 * np.abs(yr).max()
    assert np.array_equal(np.argmax(yd * yd, 1), np.argmax(yr * yr, 1))


def test_cached_environment_reuse_is_bit_exact(ctx, oracle, pkg):
    """The sweep takes the unlabelled forward factor from the environment slot of that site when the slot's provenance
    (core version, neighbour slot version, direction) says it is still current.  Same kernel, same operands: the whole
    trajectory must be bit-identical to recomputing the factor at every bond (MPST_NO_ENV_REUSE)."""
    N, T, d, C = 300, 7, 6, 2
    Xs, phi, ys, counts, cores = _problem(oracle, N, T, d, C)
    out = []
    for no_reuse in (False, True):
        ctx.debug_set("NO_ENV_REUSE", int(no_reuse))
        try:
            ctx.train_load_x(Xs, counts, d, 12)
            ctx.set_cores(cores)
            lo, gn, chi = ctx.sweep(pkg.make_opts(chi_max=12, eta=0.05), 2)
            assert ctx.debug_get("fwd_path") == (2 if no_reuse else 1)
        finally:
            ctx.debug_set("NO_ENV_REUSE", 0)
        out.append((lo, gn, chi, ctx.get_cores()))
    (l0, g0, c0, k0), (l1, g1, c1, k1) = out
    assert np.array_equal(l0, l1) and np.array_equal(g0, g1) and np.array_equal(c0, c1)
    assert all(np.array_equal(a, b) for a, b in zip(k0, k1))
    # a core replaced from the host invalidates the slot: the next bond step must not use the stale factor
    ctx.train_load_x(Xs, counts, d, 12)
    ctx.set_cores(cores)
    ctx.build_env(True)
    l_a, g_a, _ = ctx.bond_step(T - 2, True, pkg.make_opts(chi_max=12, eta=0.05))
    ctx.set_cores(cores)
    ctx.build_env(True)
    ctx.set_core(T - 2, 0.5 * cores[T - 2])                      # same shape, different values; slot T-2 is now stale
    l_b, g_b, _ = ctx.bond_step(T - 2, True, pkg.make_opts(chi_max=12, eta=0.05))
    assert ctx.debug_get("fwd_path") == 2                        # the stale slot was not used as the forward factor
    cs = [c.copy() for c in cores]
    cs[T - 2] = 0.5 * cores[T - 2]
    assert l_b != l_a
    rec = []
    oracle.fit_sweeps(cs, phi, counts, nsweeps=1, chi_max=12, eta=0.05, record=rec, max_bonds=1)
    assert abs(l_b - rec[0]["loss"]) <= 1e-9 * abs(rec[0]["loss"])


def test_phi_mode_equals_x_mode(ctx, oracle, pkg):
    """precomputed-phi entry (custom / data-driven bases) gives the same sweep as the on-device encoder."""
    Xs, phi, ys, counts, cores = _problem(oracle, 150, 6, 3)
    opts = pkg.make_opts(chi_max=6, eta=0.05)
    ctx.train_load_x(Xs, counts, 3, 6)
    ctx.set_cores(cores)
    a = ctx.sweep(opts, 1)
    ca = ctx.get_cores()
    ctx.train_load_phi(phi, counts, 6)
    ctx.set_cores(cores)
    b = ctx.sweep(opts, 1)
    cb = ctx.get_cores()
    assert np.abs(a[0][:4] - b[0][:4]).max() < 1e-9 * np.abs(a[0]).max() and np.array_equal(a[2], b[2])
    yh_b, _ = ctx.overlaps(phi_NTd=phi)
    assert np.abs(yh_b - oracle.overlaps(cb, phi)).max() < 1e-12 * np.abs(yh_b).max()


def test_api_fitMPS_classify(pkg, oracle):
    """reference-facing API: fitMPS -> classify (test/classification.jl:13-23 analogue on synthetic data):
    classify on raw X equals argmax of the oracle's overlaps with the trained cores, training accuracy
    improves, info dict carries the reference's keys."""
    def easy(n, seed):                                     # two well separated classes: rising vs falling ramps + noise
        r = np.random.default_rng(seed)
        y = r.integers(0, 2, n)
        t = np.linspace(-1, 1, 16)
        X = np.where(y[:, None] == 0, 1.0, -1.0) * t[None, :] + 0.15 * r.standard_normal((n, 16))
        return X, y
    X, y = easy(300, 11)
    Xt, yt = easy(120, 12)
    opts = pkg.MPSOptions(d=4, chi_max=10, nsweeps=3, eta=0.05, verbosity=-1)
    mps, info, test_states = pkg.fitMPS(X, y, Xt, yt, opts)
    for k in ("train_loss", "train_acc", "test_loss", "test_acc", "time_taken", "train_KL_div", "test_KL_div", "test_conf"):
        assert k in info and len(info[k]) == opts.nsweeps + 2
    assert info["train_KL_div"][-1] < info["train_KL_div"][0]
    assert info["train_acc"][-1] >= 0.95 and info["test_acc"][-1] >= 0.9
    preds = pkg.classify(mps, Xt)
    # oracle on the same trained cores
    Xs_tr, norms = oracle.transform_train_data(mps.train_data.original_data.T)
    Xs_te, _ = oracle.transform_test_data(Xt.T, norms)
    phi_te = oracle.encode(Xs_te.T, 4)
    ref = oracle.classify(mps.mps, phi_te, classes=mps.classes)
    assert np.array_equal(preds, ref)
    assert np.mean(preds == yt) == pytest.approx(info["test_acc"][-1])
    assert abs(oracle._norm2_general(mps.mps) - 1.0) < 1e-10
    clf = pkg.MPSClassifier(d=3, chi_max=6, nsweeps=2).fit(X, y)
    assert clf.predict(Xt).shape == (120,)


def test_device_metrics_match_host_reduction(ctx, oracle, pkg):
    """f1: MSE_loss_acc_conf (summary.jl:60-114) reduced on the device (mpst_eval_metrics) == the oracle's per-sample
    formulas, for a host data set with explicit labels and for the resident training set (labels = sorted class ranges),
    in x mode and phi mode; more samples than one 65 536-sample batch is covered by the chunked N below."""
    N, T, d = 70001, 6, 3                                  # two batches, ragged tail
    Xs, phi, ys, counts, cores = _problem(oracle, N, T, d)
    lab = np.repeat(np.arange(len(counts)), counts)
    ctx.train_load_x(Xs, counts, d, 8)
    ctx.set_cores(cores)
    yh, am = ctx.overlaps(X_TxN=Xs)
    onehot = np.eye(len(counts))[lab]
    mse = 0.5 * np.sum((yh - onehot) ** 2)
    kld = np.sum(-np.log(yh[np.arange(N), lab] ** 2))
    pred = np.argmax(np.abs(yh), axis=1)
    conf = np.zeros((2, 2), dtype=np.int64)
    np.add.at(conf, (lab, pred), 1)
    for kw in (dict(), dict(X_TxN=Xs, labels=lab)):
        sums, cf, n = ctx.eval_metrics(**kw)
        assert n == N and np.array_equal(cf, conf)
        assert abs(sums[0] - mse) < 1e-11 * mse and abs(sums[1] - kld) < 1e-11 * abs(kld) and sums[2] == np.sum(pred == lab)
    m_o, k_o, a_o = oracle.mse_loss_acc(cores, phi[:2000], lab[:2000])
    sums, cf, n = ctx.eval_metrics(X_TxN=Xs[:, :2000], labels=lab[:2000])
    assert abs(sums[0] / n - m_o) < 1e-11 * m_o and abs(sums[1] / n - k_o) < 1e-11 * abs(k_o) and abs(sums[2] / n - a_o) < 1e-15
    # shuffled labels on a host set: the confusion matrix follows the labels, not the sort order
    rng = np.random.default_rng(0)
    perm = rng.permutation(3000)
    sums_p, cf_p, _ = ctx.eval_metrics(X_TxN=Xs[:, perm], labels=lab[perm])
    sums_s, cf_s, _ = ctx.eval_metrics(X_TxN=Xs[:, np.sort(perm)], labels=lab[np.sort(perm)])
    assert np.array_equal(cf_p, cf_s) and abs(sums_p[0] - sums_s[0]) < 1e-11 * sums_s[0]
    ctx.train_load_phi(phi[:5000], np.array([np.sum(lab[:5000] == 0), np.sum(lab[:5000] == 1)]), 8)
    ctx.set_cores(cores)
    s1, c1, n1 = ctx.eval_metrics()
    s2, c2, n2 = ctx.eval_metrics(phi_NTd=phi[:5000], labels=lab[:5000])
    assert n1 == n2 == 5000 and np.array_equal(c1, c2) and np.allclose(s1, s2, rtol=1e-13)
