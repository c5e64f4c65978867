"""f3: data-driven / time-dependent encodings evaluated on the device from per-site coefficient tables
(csrc/encode_table.cu) against the oracle's restatement of the reference's per-point encoders (bases.jl:95-129,
splitbases.jl:96-163), and through training / classify (== the precomputed-phi path fed by the oracle)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(5)
    T, N = 7, 600
    t = np.arange(1, T + 1)
    X = np.sin(2 * np.pi * t[:, None] / 5.0 + rng.uniform(0, 2 * np.pi, N)[None, :]) * 0.7 + 0.1 * rng.standard_normal((T, N))
    return np.clip(X, -1.0, 1.0)


def test_projected_legendre_table(ctx, oracle, pkg, data):
    from mpstime_jl_b200 import encodings_host as eh
    d = 6
    for norm in (False, True):
        (kind, ns, ip, dp), orders = eh.project_legendre(data, d, norm=norm)
        assert orders.shape == (data.shape[0], d) and 1 <= orders.min() and orders.max() <= 7 * d and len(set(orders[0])) == d
        ctx.set_encoding_table(kind, ns, d, ip, dp)
        ctx.model_init(data.shape[0], 2, d, 4, basis=kind)
        x = np.linspace(-1, 1, 1001)
        for site in (0, 3, data.shape[0] - 1):
            dev = ctx.encode_site(site, x)
            ref = oracle.projected_legendre_encode(x, orders[site], norm)
            assert np.abs(dev - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_sahand_legendre_table(ctx, oracle, pkg, data):
    from mpstime_jl_b200 import encodings_host as eh
    d = 5
    for td in (True, False):
        kind, ns, ip, dp = eh.init_sahand_legendre(data, d, time_dependent=td)
        assert ns == (data.shape[0] if td else 1)
        ctx.set_encoding_table(kind, ns, d, ip, dp)
        ctx.model_init(data.shape[0], 2, d, 4, basis=kind)
        x = np.concatenate([np.linspace(-1, 1, 777), [-1.0, 1.0, 0.0]])
        for site in ((0, 2, 6) if td else (0, 5)):
            s = site if td else 0
            npts = ip[s, 0]
            kde_x = dp[s, 0] + dp[s, 1] * np.arange(npts)
            ref = oracle.sahand_legendre_encode(x, d, kde_x, dp[s, 4 + d * d:], dp[s, 2], dp[s, 3], dp[s, 4:4 + d * d].reshape(d, d))
            dev = ctx.encode_site(site, x)
            assert np.abs(dev - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_split_table(ctx, oracle, pkg, data):
    from mpstime_jl_b200 import encodings_host as eh
    X01 = (data + 1.0) / 2.0
    for method, aux, ad, d in (("hist", "uniform", 2, 8), ("unif", "uniform", 1, 5), ("hist", "legendre", 3, 6)):
        rng_ = (0.0, 1.0) if aux == "uniform" else (-1.0, 1.0)
        Xn = X01 if aux == "uniform" else data
        kind, ns, ip, dp = eh.split_table(Xn, d, ad, aux=aux, method=method, enc_range=rng_)
        ctx.set_encoding_table(kind, ns, d, ip, dp)
        ctx.model_init(Xn.shape[0], 2, d, 4, basis=kind)
        for site in (0, 4):
            bins = dp[site if ns > 1 else 0]
            x = np.concatenate([np.linspace(rng_[0], rng_[1], 501), bins])    # bin edges exactly: the 1/2-1/2 rule
            dev = ctx.encode_site(site, x)
            ref = oracle.split_encode(x, bins, ad, aux)
            assert np.abs(dev - ref).max() < 1e-13
            if aux == "uniform":
                assert np.allclose(dev.sum(axis=1), 1.0)                       # normalised encoding, edges included


@pytest.mark.parametrize("enc,kw", [("Legendre_No_Norm", dict(projected_basis=True, d=5)), ("SLTD", dict(d=4)),
                                    ("hist_split_uniform", dict(d=6, aux_basis_dim=2))])
def test_fit_and_classify_with_table_encodings(ctx, oracle, pkg, enc, kw):
    """fitMPS / classify with a data-driven encoding == the same sweep fed with precomputed phi produced by the oracle's
    encoders from the same tables (mpst_train_load_phi), bond by bond."""
    from mpstime_jl_b200 import encodings_host as eh
    X, y = oracle.synthetic_two_class(240, 9, seed=3)
    opts = pkg.MPSOptions(encoding=enc, chi_max=8, nsweeps=1, eta=0.05, verbosity=-1, log_level=0, **kw)
    name = opts._check()
    Xs, norms = pkg.transform_train_data(X.T, opts)
    Xs_sorted, Xo, ys, order, classes, counts = pkg.sort_by_class(Xs, X, y)
    kind, ns, ip, dp = pkg.api.build_encoding_table(opts, Xs_sorted)
    d, T = opts.d, X.shape[1]
    # oracle phi (N, T, d) from the same tables
    phi = np.empty((Xs_sorted.shape[1], T, d))
    for j in range(T):
        s = j if ns > 1 else 0
        if name == "table_legendre_proj":
            phi[:, j] = oracle.projected_legendre_encode(Xs_sorted[j], ip[s, :d], False)
        elif name == "table_sahand_legendre":
            kx = dp[s, 0] + dp[s, 1] * np.arange(ip[s, 0])
            phi[:, j] = oracle.sahand_legendre_encode(Xs_sorted[j], d, kx, dp[s, 4 + d * d:], dp[s, 2], dp[s, 3], dp[s, 4:4 + d * d].reshape(d, d)) if ip[s, 1] else 0.0
        else:
            phi[:, j] = oracle.split_encode(Xs_sorted[j], dp[s], ip[s, 1], "uniform")
    cores = pkg.generate_starting_mps(4, T, d, 2, seed=1234)
    topts = pkg.make_opts(chi_max=8, eta=0.05)
    ctx.set_encoding_table(kind, ns, d, ip, dp)
    ctx.train_load_x(Xs_sorted, counts, d, 8, basis=kind)
    ctx.set_cores(cores)
    a = ctx.sweep(topts, 1)
    ca = ctx.get_cores()
    ctx.train_load_phi(phi, counts, 8)
    ctx.set_cores(cores)
    b = ctx.sweep(topts, 1)
    assert np.array_equal(a[2], b[2]) and np.abs(a[0][:4] - b[0][:4]).max() < 1e-9 * np.abs(b[0]).max()
    # oracle bond 1 from the oracle's own phi
    rec = []
    oracle.fit_sweeps(cores, phi, counts, nsweeps=1, chi_max=8, eta=0.05, record=rec, max_bonds=2)
    assert abs(a[0][0] - rec[0]["loss"]) < 1e-9 * abs(rec[0]["loss"]) and abs(a[1][1] - rec[1]["gradnorm"]) < 1e-8 * rec[1]["gradnorm"]
    # public API: fitMPS keeps the table in the TrainedMPS, classify re-uses it
    mps, info, _ = pkg.fitMPS(X, y, opts=opts)
    assert mps.enc_table is not None and mps.enc_table[0] == kind
    pred = pkg.classify(mps, X)
    ctx2 = ctx
    ctx2.set_encoding_table(kind, ns, d, ip, dp)
    ctx2.model_init(T, 2, d, 8, basis=kind)
    ctx2.set_cores(mps.mps)
    Xt, _ = pkg.transform_test_data(X.T, norms, opts)
    yh, am = ctx2.overlaps(X_TxN=Xt)
    assert np.array_equal(pred, np.asarray(mps.classes)[am])
    # imputation runs with the same tables (K8 table mode; parity in tests/test_gpu_zz_impute_tables.py)
    imp = pkg.init_imputation_problem(mps, X, y, dx=1e-3, verbosity=-1)
    res = pkg.MPS_impute(imp, int(classes[0]), 0, [2, 3], "mode")
    ts = np.asarray(res[0])
    assert ts.shape[-1] == T and np.isfinite(ts).all()
