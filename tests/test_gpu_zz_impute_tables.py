"""K8 with the data-driven / time-dependent encodings (projected Legendre, SLTD, histogram-split bases): the reference
pre-computes one set of encoded grid states PER SITE (`xvals_enc[site]`, Imputation/imputation.jl:92-100) and encodes
known / imputed values with the site's own basis (`get_state(x, opts, j, enc_args)`).  The device does the same from the
per-site coefficient tables (impute.cu `impute_kernel<true>`); the oracle replays `impute_at!` (MPS_methods.jl:93-180)
with per-site encoders built from the same tables.  Tolerance 1e-8 on imputed values (BASELINE.json north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PATTERNS = [[3, 4, 5, 6], [0, 1], [8, 9], [1, 3, 6], [5]]


def _site_encoder(oracle, name, d, ns, ip, dp):
    def enc(j, x):
        s = j if ns > 1 else 0
        x = np.asarray(x, dtype=np.float64)
        if name == "table_legendre_proj":
            return oracle.projected_legendre_encode(x, ip[s, :d], False)
        if name == "table_sahand_legendre":
            if not ip[s, 1]:
                return np.zeros(x.shape + (d,))
            kx = dp[s, 0] + dp[s, 1] * np.arange(ip[s, 0])
            return oracle.sahand_legendre_encode(x, d, kx, dp[s, 4 + d * d:], dp[s, 2], dp[s, 3], dp[s, 4:4 + d * d].reshape(d, d))
        return oracle.split_encode(x, dp[s], ip[s, 1], "uniform")
    return enc


@pytest.mark.parametrize("enc,kw,methods", [
    ("SLTD", dict(d=4), ("median", "mode", "mean", "ITS")),
    # the projected orders of this toy set are all odd at some sites: p(x) = p(-x), the median is exactly 0 and the state
    # of x = 0 vanishes (in the reference too), so only the methods that do not land on 0 are compared
    ("Legendre_No_Norm", dict(projected_basis=True, d=5), ("ITS", "mode")),
    # piecewise-constant pdf: the mode is a whole bin (ties), so only the cdf-based methods are compared
    ("hist_split_uniform", dict(d=6, aux_basis_dim=2), ("median", "ITS")),
])
@pytest.mark.parametrize("order", ["forwards", "backwards"])
def test_impute_with_table_encodings(ctx, oracle, pkg, enc, kw, methods, order):
    X, y = oracle.synthetic_two_class(240, 10, seed=3)
    opts = pkg.MPSOptions(encoding=enc, chi_max=8, nsweeps=2, eta=0.05, verbosity=-1, log_level=0, **kw)
    name = opts._check()
    Xs, norms = pkg.transform_train_data(X.T, opts)
    Xs_sorted, Xo, ys, perm, classes, counts = pkg.sort_by_class(Xs, X, y)
    kind, ns, ip, dp = pkg.api.build_encoding_table(opts, Xs_sorted)
    d, T = opts.d, X.shape[1]
    site_enc = _site_encoder(oracle, name, d, ns, ip, dp)
    ctx.set_encoding_table(kind, ns, d, ip, dp)
    ctx.train_load_x(Xs_sorted, counts, d, 8, basis=kind)
    ctx.set_cores(pkg.generate_starting_mps(4, T, d, 2, seed=1234))
    ctx.sweep(pkg.make_opts(chi_max=8, eta=0.05), 2)
    cores = ctx.get_cores()
    grid = oracle.make_grid(pkg.api.encoding_range(opts.encoding), 1e-3)
    genc = np.stack([site_enc(j, grid) for j in range(T)])                  # (T, G, d)
    rng = np.random.default_rng(1)
    n = len(PATTERNS)
    for cls in range(2):
        class_cores = oracle.expand_label_index(cores)[cls]
        Xb = Xs_sorted[:, rng.integers(0, Xs_sorted.shape[1], n)].copy()
        mask = np.zeros((T, n), dtype=np.uint8)
        for k, ms in enumerate(PATTERNS):
            mask[ms, k] = 1
            Xb[ms, k] = grid[len(grid) // 3]
        Kmax = int(mask.sum(0).max())
        U = rng.uniform(0.05, 0.95, size=(n, 1, Kmax))
        for method in methods:
            out = ctx.impute_batch(cls, Xb, mask, grid, method=method, uniforms=U if method == "ITS" else None,
                                   impute_order=order)
            for k, ms in enumerate(PATTERNS):
                ref, _, _, _ = oracle.impute_series_ex(class_cores, Xb[:, k], ms, grid, genc, d, basis=site_enc, method=method,
                                                       uniforms=U[k, 0] if method == "ITS" else None, impute_order=order)
                dev = out[k, 0]
                known = np.setdiff1d(np.arange(T), ms)
                assert np.array_equal(dev[known], Xb[known, k])
                assert np.abs(dev[ms] - ref[ms]).max() < 1e-8, (enc, cls, method, k, dev[ms], ref[ms])
