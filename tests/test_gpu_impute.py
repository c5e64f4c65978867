"""K8 parity: batched device imputation vs the oracle's restatement of MPS_impute (median / mode / mean /
ITS), through the C ABI.  Tolerance 1e-8 on imputed values (BASELINE.json north_star); grid-snapped methods
must pick the same grid point."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trained(oracle):
    N, T, d, C = 120, 12, 4, 2
    X, y = oracle.synthetic_two_class(N, T, seed=21)
    Xs, norms = oracle.transform_train_data(X.T)
    phi, ys, order, counts, classes = oracle.encode_dataset(Xs, y, d)
    cores = oracle.random_start_mps(T, d, 4, C, seed=3)
    cores = oracle.fit_sweeps(cores, phi, counts, nsweeps=3, chi_max=9, eta=0.05)
    return dict(cores=cores, Xs=Xs[:, order], T=T, d=d, C=C)


PATTERNS = [
    [4, 5, 6, 7],            # contiguous block in the bulk
    [0, 1, 2],               # block at the left edge (no LE)
    [9, 10, 11],             # block at the right edge (no known site to the right)
    [1, 3, 5, 8],            # scattered: known runs between missing sites
    [6],                     # single site
    list(range(12)),         # everything missing
    [0, 11],                 # both ends
]


@pytest.mark.parametrize("method", ["median", "mode", "mean", "ITS"])
def test_impute_batch_matches_oracle(ctx, oracle, trained, method):
    T, d, C = trained["T"], trained["d"], trained["C"]
    cores = trained["cores"]
    grid = oracle.make_grid((-1.0, 1.0), 1e-3)              # 2001 points: seconds on the CPU oracle
    genc = oracle.encode(grid, d)
    chi = max(max(A.shape[0], A.shape[2]) for A in cores)
    ctx.model_init(T, C, d, chi)
    ctx.set_cores(cores)
    rng = np.random.default_rng(0)
    for cls in range(C):
        class_cores = oracle.expand_label_index(cores)[cls]
        n = len(PATTERNS)
        X = trained["Xs"][:, rng.integers(0, trained["Xs"].shape[1], n)].copy()      # (T, n)
        mask = np.zeros((T, n), dtype=np.uint8)
        for k, ms in enumerate(PATTERNS):
            mask[ms, k] = 1
            X[ms, k] = 0.123                                                         # the fill value is irrelevant
        Kmax = int(mask.sum(0).max())
        U = rng.uniform(0.02, 0.98, size=(n, 2, Kmax))
        out = ctx.impute_batch(cls, X, mask, grid, method=method, uniforms=U if method == "ITS" else None,
                               n_traj=2 if method == "ITS" else 1)
        assert out.shape == (n, 2 if method == "ITS" else 1, T)
        for k, ms in enumerate(PATTERNS):
            for tr in range(out.shape[1]):
                ref, idx = oracle.impute_series(class_cores, X[:, k], ms, grid, genc, d, method=method,
                                                uniforms=U[k, tr] if method == "ITS" else None)
                dev = out[k, tr]
                known = np.setdiff1d(np.arange(T), ms)
                assert np.array_equal(dev[known], X[known, k])
                assert np.abs(dev[ms] - ref[ms]).max() < 1e-8, (cls, k, tr, dev[ms], ref[ms])


def test_impute_mode_max_jump_and_no_missing(ctx, oracle, trained):
    T, d, C = trained["T"], trained["d"], trained["C"]
    cores = trained["cores"]
    grid = oracle.make_grid((-1.0, 1.0), 1e-3)
    genc = oracle.encode(grid, d)
    ctx.model_init(T, C, d, max(max(A.shape[0], A.shape[2]) for A in cores))
    ctx.set_cores(cores)
    X = trained["Xs"][:, :3].copy()
    mask = np.zeros((T, 3), dtype=np.uint8)
    mask[4:8, 0] = 1
    mask[2:5, 1] = 1                                       # instance 2 has nothing missing
    out = ctx.impute_batch(0, X, mask, grid, method="mode", max_jump=0.05)
    cls = oracle.expand_label_index(cores)[0]
    for k, ms in ((0, [4, 5, 6, 7]), (1, [2, 3, 4])):
        ref, _ = oracle.impute_series(cls, X[:, k], ms, grid, genc, d, method="mode", max_jump=0.05)
        assert np.abs(out[k, 0, ms] - ref[ms]).max() < 1e-8
        assert np.all(np.abs(np.diff(np.concatenate([[X[ms[0] - 1, k]], out[k, 0, ms]]))) <= 0.05 + 1e-12)
    assert np.array_equal(out[2, 0], X[:, 2])


def test_api_mps_impute(pkg, oracle):
    """reference-facing path: fitMPS -> init_imputation_problem -> MPS_impute, against the oracle run on the
    same trained cores (normalisation, mean fill, inverse transform included)."""
    X, y = oracle.synthetic_two_class(200, 20, seed=5)
    Xt, yt = oracle.synthetic_two_class(20, 20, seed=6)
    opts = pkg.MPSOptions(d=4, chi_max=8, nsweeps=2, eta=0.05, verbosity=-1, log_level=0)
    mps, info, _ = pkg.fitMPS(X, y, opts=opts)
    imp = pkg.init_imputation_problem(mps, Xt, yt, dx=1e-3)
    ms = list(range(8, 14))
    ts, err, target, stats = pkg.MPS_impute(imp, 1, 3, ms, "median")
    assert len(ts) == 1 and ts[0].shape == (20,) and "MAE" in stats[0]
    # oracle replay
    Xs_tr, norms = oracle.transform_train_data(mps.train_data.original_data.T)
    raw = Xt[np.nonzero(yt == 1)[0][3]].copy()
    filled = raw.copy()
    filled[ms] = np.mean(mps.train_data.original_data)
    xs, oob = oracle.transform_test_data(filled, norms)
    grid = oracle.make_grid((-1.0, 1.0), 1e-3)
    cls = oracle.expand_label_index(mps.mps)[1]
    ref, _ = oracle.impute_series(cls, xs, ms, grid, oracle.encode(grid, 4), 4)
    ref_raw = oracle.invert_test_transform(ref, oob, norms)
    assert np.abs(ts[0][ms] - ref_raw[ms]).max() < 1e-8 * max(1.0, np.abs(ref_raw).max())
    known = np.setdiff1d(np.arange(20), ms)
    assert np.abs(ts[0][known] - raw[known]).max() < 1e-9
    assert np.array_equal(target, raw)


@pytest.mark.parametrize("method,order", [("median", "forwards"), ("median", "backwards"), ("mean", "forwards"),
                                          ("mean", "backwards"), ("mode", "backwards"), ("ITS", "backwards")])
def test_impute_error_bars_and_backwards_order(ctx, oracle, trained, method, order):
    """a18: WMAD error bars of the median (sampling_utils.jl:192-196), the mean's standard deviation (:89-97) and
    impute_order = :backwards (MPS_methods.jl:113-118; on the device: the forward walk on the mirrored chain) against
    the oracle's literal restatement (QR sweep to the last site, right-to-left walk)."""
    T, d, C = trained["T"], trained["d"], trained["C"]
    cores = trained["cores"]
    grid = oracle.make_grid((-1.0, 1.0), 1e-3)
    genc = oracle.encode(grid, d)
    ctx.model_init(T, C, d, max(max(A.shape[0], A.shape[2]) for A in cores))
    ctx.set_cores(cores)
    rng = np.random.default_rng(3)
    n = len(PATTERNS)
    X = trained["Xs"][:, rng.integers(0, trained["Xs"].shape[1], n)].copy()
    mask = np.zeros((T, n), dtype=np.uint8)
    for k, ms in enumerate(PATTERNS):
        mask[ms, k] = 1
        X[ms, k] = -0.3
    Kmax = int(mask.sum(0).max())
    U = rng.uniform(0.02, 0.98, size=(n, 1, Kmax))
    get_err = method in ("median", "mean")
    out, err = ctx.impute_batch(1, X, mask, grid, method=method, uniforms=U if method == "ITS" else None, get_err=get_err,
                                impute_order=order, max_jump=0.4 if method == "mode" else -1.0, return_err=True)
    cls = oracle.expand_label_index(cores)[1]
    for k, ms in enumerate(PATTERNS):
        ref, _, eref, _ = oracle.impute_series_ex(cls, X[:, k], ms, grid, genc, d, method=method,
                                                  uniforms=U[k, 0] if method == "ITS" else None, impute_order=order,
                                                  get_err=get_err, max_jump=0.4 if method == "mode" else None)
        assert np.abs(out[k, 0, ms] - ref[ms]).max() < 1e-8, (k, out[k, 0, ms], ref[ms])
        known = np.setdiff1d(np.arange(T), ms)
        assert np.array_equal(out[k, 0, known], X[known, k]) and np.all(err[k, 0, known] == 0.0)
        assert np.abs(err[k, 0, ms] - eref[ms]).max() < 1e-8, (k, err[k, 0, ms], eref[ms])
        if get_err:
            assert np.all(err[k, 0, ms] > 0.0)


@pytest.mark.parametrize("order", ["forwards", "backwards"])
def test_its_rejection_sampling(ctx, oracle, trained, order):
    """a18: ITS with WMAD rejection (sampling_utils.jl:291-311): the flat uniform stream is consumed in order over sites
    and trajectories, a draw is accepted when |x - median| < threshold * WMAD, the last of max_trials draws stands."""
    T, d, C = trained["T"], trained["d"], trained["C"]
    cores = trained["cores"]
    grid = oracle.make_grid((-1.0, 1.0), 1e-3)
    genc = oracle.encode(grid, d)
    ctx.model_init(T, C, d, max(max(A.shape[0], A.shape[2]) for A in cores))
    ctx.set_cores(cores)
    rng = np.random.default_rng(9)
    n, ntraj, max_trials, thr = len(PATTERNS), 2, 4, 1.0
    X = trained["Xs"][:, rng.integers(0, trained["Xs"].shape[1], n)].copy()
    mask = np.zeros((T, n), dtype=np.uint8)
    for k, ms in enumerate(PATTERNS):
        mask[ms, k] = 1
    Kmax = int(mask.sum(0).max())
    U = rng.uniform(0.0, 1.0, size=(n, ntraj * Kmax * max_trials))
    out, err = ctx.impute_batch(0, X, mask, grid, method="ITS", uniforms=U, n_traj=ntraj, impute_order=order,
                                rejection_threshold=thr, max_trials=max_trials, return_err=True)
    cls = oracle.expand_label_index(cores)[0]
    used_more_than_one = False
    for k, ms in enumerate(PATTERNS):
        cur = 0
        for tr in range(ntraj):
            ref, _, eref, used = oracle.impute_series_ex(cls, X[:, k], ms, grid, genc, d, method="ITS", uniforms=U[k, cur:],
                                                         impute_order=order, rejection_threshold=thr, max_trials=max_trials)
            cur += used
            used_more_than_one |= used > len(ms)
            assert np.abs(out[k, tr, ms] - ref[ms]).max() < 1e-8, (k, tr)
            assert np.abs(err[k, tr, ms] - eref[ms]).max() < 1e-8, (k, tr)
    assert used_more_than_one, "no draw was ever rejected: the test does not exercise the rejection loop"


def test_api_mps_impute_error_bars(pkg, oracle):
    """MPS_impute(..., :median; get_wmad=true) through the reference-facing API: pred_err is transformed back the way
    get_predictions does (imputation.jl:337-384)."""
    X, y = oracle.synthetic_two_class(200, 20, seed=5)
    Xt, yt = oracle.synthetic_two_class(20, 20, seed=6)
    opts = pkg.MPSOptions(d=4, chi_max=8, nsweeps=2, eta=0.05, verbosity=-1, log_level=0)
    mps, info, _ = pkg.fitMPS(X, y, opts=opts)
    imp = pkg.init_imputation_problem(mps, Xt, yt, dx=1e-3)
    ms = list(range(8, 14))
    ts, err, target, stats = pkg.MPS_impute(imp, 1, 3, ms, "median", get_wmad=True)
    assert err[0].shape == (20,) and np.all(np.isfinite(err[0][ms])) and np.all(err[0][ms] > 0)
    ts2, err2, _, _ = pkg.MPS_impute(imp, 1, 3, ms, "median")
    assert np.array_equal(ts[0], ts2[0]) and np.all(err2[0][ms] == 0.0)            # get_wmad defaults to false (:210)
    assert pkg.MPS_impute(imp, 1, 3, ms, "ITS", num_trajectories=2)[1] == [None, None]
    # oracle replay of the error bars in normalised space
    Xs_tr, norms = oracle.transform_train_data(mps.train_data.original_data.T)
    raw = Xt[np.nonzero(yt == 1)[0][3]].copy()
    filled = raw.copy()
    filled[ms] = np.mean(mps.train_data.original_data)
    xs, oob = oracle.transform_test_data(filled, norms)
    grid = oracle.make_grid((-1.0, 1.0), 1e-3)
    cls = oracle.expand_label_index(mps.mps)[1]
    ref, _, eref, _ = oracle.impute_series_ex(cls, xs, ms, grid, oracle.encode(grid, 4), 4, get_err=True)
    back = oracle.invert_test_transform(ref + eref, oob, norms) - oracle.invert_test_transform(ref, oob, norms)
    assert np.abs(err[0][ms] - back[ms]).max() < 1e-7 * max(1.0, np.abs(back[ms]).max())
    with pytest.raises(ValueError):
        pkg.MPS_impute(imp, 1, 3, ms, "median", impute_order="sideways")
