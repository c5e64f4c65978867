"""Host-side data preparation that stays on the CPU in the reference as well: normalisation
(reference src/utils.jl:161-334), class sort (Encodings/encodings.jl:33-46) and the random
starting MPS (Training/RealRealHighDimension.jl:1-41).  None of this is on the hot path."""
import numpy as np

from .core import BASIS_IDS, BASIS_RANGE


_SL_NAMES = ("sl", "sahand_legendre", "sahand_legendre_time_independent", "sahand-legendre_time_independent",
             "sltd", "sahand_legendre_time_dependent", "sahand-_legendre_time_dependent")


def split_encoding_name(basis):
    """("hist" | "unif", auxiliary basis name) of a split-basis symbol (options.jl:261-272), else None."""
    s = str(basis).lower().lstrip(":")
    for pre, kind in (("hist_split_", "hist"), ("hist._split_", "hist"), ("histogram_split_", "hist"),
                      ("unif_split_", "unif"), ("unif._split_", "unif"), ("uniform_split_", "unif")):
        if s.startswith(pre):
            return kind, s[len(pre):]
    return None


def encoding_range(basis):
    """Domain of a basis; accepts the reference's spellings ("Legendre_No_Norm", ":legendre", ":SLTD",
    ":hist_split_uniform", ...) -- model_encoding, Structs/options.jl:243-279."""
    s = str(basis).lower().lstrip(":")
    if s in _SL_NAMES:
        return (-1.0, 1.0)
    sp = split_encoding_name(s)
    if sp is not None:
        return encoding_range(sp[1])                          # a split basis inherits the range of its auxiliary basis
    return BASIS_RANGE[BASIS_IDS[s]]


class Norms:
    """[sig_trans, minmax] of transform_train_data (utils.jl:161-200): global RobustSigmoid
    (median, IQR/1.35; options.jl:72-77) followed by global MinMax."""

    def __init__(self, sigmoid=None, minmax=None):
        self.sigmoid = sigmoid
        self.minmax = minmax

    def apply(self, X):
        X = np.array(X, dtype=np.float64, copy=True)
        if self.sigmoid is not None:
            m, iqr = self.sigmoid
            X = 1.0 / (1.0 + np.exp(-(X - m) / (iqr / 1.35)))
        if self.minmax is not None:
            lo, hi = self.minmax
            X = (X - lo) / (hi - lo)
        return X

    def invert(self, X):
        X = np.array(X, dtype=np.float64, copy=True)
        if self.minmax is not None:
            lo, hi = self.minmax
            X = X * (hi - lo) + lo
        if self.sigmoid is not None:
            m, iqr = self.sigmoid
            with np.errstate(invalid="ignore", divide="ignore"):
                X = np.where((X > 0) & (X < 1), np.log(X / (1.0 - X)), np.nan) * (iqr / 1.35) + m
        return X


def transform_train_data(X, opts):
    """utils.jl:161-200.  X: (T, N) series as columns (statistics are global)."""
    X = np.asarray(X, dtype=np.float64)
    norms = Norms()
    if opts.sigmoid_transform:
        q25, med, q75 = np.quantile(X, [0.25, 0.5, 0.75])      # one partition pass; quantile(0.5) == median
        norms.sigmoid = (float(med), float(q75 - q25))
    Xs = Norms(norms.sigmoid, None).apply(X)
    if opts.minmax:
        norms.minmax = (float(Xs.min()), float(Xs.max()))
        lo, hi = norms.minmax
        Xs = (Xs - lo) / (hi - lo)
        lb, ub = opts.data_bounds
        Xs = Xs * (ub - lb) + lb
    a, b = encoding_range(opts.encoding)
    return (b - a) * Xs + a, norms


def transform_test_data(X, norms, opts, rescale_out_of_bounds=True):
    """utils.jl:202-278.  X: (T, n) or (T,).  Returns (X_scaled, oob_rescales)."""
    X = np.asarray(X, dtype=np.float64)
    single = X.ndim == 1
    Xs = X.reshape(-1, 1) if single else X
    if Xs.size == 0:
        return Xs.copy(), []
    Xs = norms.apply(Xs)
    if opts.minmax:
        lb, ub = opts.data_bounds
        Xs = Xs * (ub - lb) + lb
    oob = []
    if rescale_out_of_bounds:
        # per series (utils.jl:236-262): shift up when the minimum is below 0, then divide when the maximum is above 1;
        # all series at once -- the same element-wise operations as the per-column loop, so the same bits
        Xs = np.ascontiguousarray(Xs)
        lo = Xs.min(axis=0)
        low = lo < 0
        if low.any():
            Xs[:, low] -= lo[low]
        hi = Xs.max(axis=0)
        high = hi > 1
        if high.any():
            Xs[:, high] /= hi[high]
        for i in np.nonzero(low | high)[0]:
            oob.append((int(i), float(lo[i]) if low[i] else 0.0, float(hi[i]) if high[i] else 1.0))
    a, b = encoding_range(opts.encoding)
    Xs = (b - a) * Xs + a
    return (Xs[:, 0] if single else Xs), oob


def invert_test_transform(Xs, oob, norms, opts):
    """utils.jl:299-330."""
    Xs = np.asarray(Xs, dtype=np.float64)
    single = Xs.ndim == 1
    X = Xs.reshape(-1, 1).copy() if single else Xs.copy()
    a, b = encoding_range(opts.encoding)
    X = (X - a) / (b - a)
    for (i, lb_s, ub_s) in oob:
        X[:, i] = X[:, i] * ub_s + lb_s
    if opts.minmax:
        lb, ub = opts.data_bounds
        X = (X - lb) / (ub - lb)
    X = norms.invert(X)
    return X[:, 0] if single else X


def sort_by_class(X_scaled_TxN, X_orig_NxT, y):
    """encode_dataset (encodings.jl:33-46): stable sortperm(y); class_distribution (:151-152)."""
    y = np.asarray(y)
    order = np.argsort(y, kind="stable")
    ys = y[order]
    classes, counts = np.unique(ys, return_counts=True)
    return X_scaled_TxN[:, order], (X_orig_NxT[order] if X_orig_NxT is not None else None), ys, order, classes, counts


def generate_starting_mps(chi_init, T, d, C, seed=1234):
    """Structure of generate_startingMPS (RealRealHighDimension.jl:1-41): random MPS, uniform link
    dimension chi_init, class index on the LAST site, unit norm, orthogonality centre at the last
    site.  (ITensors' random_mps stream is not reproducible outside Julia; the Julia shim keeps
    generate_startingMPS and passes its cores through mpst_set_core.)"""
    rng = np.random.default_rng(seed)
    chis = [1] + [int(min(chi_init, d ** min(j, T - j, 30))) for j in range(1, T)] + [1]
    cores = [rng.standard_normal((chis[j], d, chis[j + 1]) + ((C,) if j == T - 1 else ())) for j in range(T)]
    for j in range(T - 1):
        a, s, b = cores[j].shape
        Q, R = np.linalg.qr(cores[j].reshape(a * s, b))
        cores[j] = Q.reshape(a, s, Q.shape[1])
        nxt = np.tensordot(R, cores[j + 1], axes=(1, 0))
        cores[j + 1] = nxt / np.linalg.norm(nxt)             # keep the running scale at 1 (long chains overflow otherwise)
    cores[-1] = cores[-1] / np.linalg.norm(cores[-1])
    return cores
