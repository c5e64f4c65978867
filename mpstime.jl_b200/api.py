"""Host-side mirror of the reference's public interface for the hot path:
    MPSOptions / fitMPS / classify / init_imputation_problem / MPS_impute / MPSClassifier
(reference src/Structs/options.jl, src/Training/RealRealHighDimension.jl:383-890, src/summary.jl:116-177,
src/Imputation/imputation.jl:48-563, src/MLJIntegration/MLJ_integration.jl).  Julia is not present in
this image, so this Python layer plays the role of the Julia shim: same names,
argument meaning and error behaviour; every numeric step of the path goes through the C ABI."""
from dataclasses import dataclass, field, replace
from typing import Optional, Tuple

import numpy as np

from . import dist as _dist
from .core import Context, MPSTError, make_opts, BASIS_IDS
from . import encodings_host as _eh
from .preprocess import (_SL_NAMES, encoding_range, generate_starting_mps, invert_test_transform, sort_by_class,
                         split_encoding_name, transform_test_data, transform_train_data)

_ENC = {"legendre_no_norm": "legendre_no_norm", "legendre": "legendre_no_norm", "legendre_norm": "legendre_norm",
        "uniform": "uniform", "fourier": "fourier", "stoudenmire": "stoudenmire", "sahand": "sahand"}


@dataclass
class MPSOptions:
    """Same 27 fields and defaults as the reference's MPSOptions (Structs/options.jl:106-134)."""
    verbosity: int = 1
    nsweeps: int = 10
    chi_max: int = 25
    eta: float = 0.01
    d: int = 5
    encoding: str = "Legendre_No_Norm"
    projected_basis: bool = False
    aux_basis_dim: int = 2
    cutoff: float = 1e-10
    update_iters: int = 1
    dtype: type = np.float64
    loss_grad: str = "KLD"
    bbopt: str = "TSGO"
    track_cost: bool = False
    rescale: Tuple[bool, bool] = (False, True)
    train_classes_separately: bool = False
    encode_classes_separately: bool = False
    return_encoding_meta_info: bool = False
    minmax: bool = True
    exit_early: bool = False
    sigmoid_transform: bool = True
    init_rng: int = 1234
    chi_init: int = 4
    log_level: int = 3
    data_bounds: Tuple[float, float] = (0.0, 1.0)
    use_legacy_ITensor: bool = False
    svd_alg: str = "divide_and_conquer"

    def _table_spec(self):
        """None for a built-in basis, else how to build the per-site coefficient table of a data-driven / time-dependent
        encoding (model_encoding, Structs/options.jl:243-279): ("legendre_proj", norm) | ("sahand_legendre",
        time_dependent) | ("split", method, aux)."""
        enc = str(self.encoding).lower().lstrip(":")
        if enc in _SL_NAMES:
            return ("sahand_legendre", enc.startswith("sltd") or "time_dependent" in enc)
        sp = split_encoding_name(enc)
        if sp is not None:
            if sp[1] not in _eh.AUX_IDS:
                raise ValueError(f"split basis over {sp[1]!r}: the auxiliary basis must be a data-independent real basis "
                                 "(Uniform, Legendre, Legendre_Norm); data-driven auxiliary bases are forbidden in the "
                                 "reference too (splitbases.jl:33)")
            if self.d % self.aux_basis_dim:
                raise ValueError(f"The auxilliary basis dimension ({self.aux_basis_dim}) must evenly divide the total "
                                 f"feature dimension ({self.d})")                          # splitbases.jl:4-6
            return ("split", sp[0], sp[1])
        if self.projected_basis and enc in ("legendre", "legendre_no_norm", "legendre_norm"):
            return ("legendre_proj", enc == "legendre_norm")
        return None

    def _check(self):
        enc = str(self.encoding).lower().lstrip(":")
        table = self._table_spec()
        if enc not in _ENC and table is None:
            raise ValueError(f"encoding {self.encoding!r}: unknown to the B200 backend (built-in: Legendre, Legendre_No_Norm, "
                             "Legendre_Norm, Uniform; from per-site tables: projected Legendre, SL / SLTD, hist_split_* / "
                             "unif_split_*; Fourier / Stoudenmire / Sahand only through mpst_encode)")
        if enc in ("fourier", "stoudenmire", "sahand"):
            # loss_functions.jl:343 keeps yhat in a Ref{Float64}: the array path is real-only
            raise ValueError("complex encodings cannot be trained on the array path (reference: InexactError)")
        if str(self.bbopt).upper().lstrip(":") not in ("TSGO", "GD"):
            # loss_functions.jl:166-170
            raise ValueError("Optim/OptimKit based solvers currently unimplemented for this version, "
                             "set 'use_legacy_ITensor=true' in MPSOptions to enable")
        if str(self.loss_grad).upper().lstrip(":") not in ("KLD", "MSE"):
            raise ValueError("loss_grad must be :KLD or :MSE on the array path")
        if self.use_legacy_ITensor:
            raise ValueError("use_legacy_ITensor=true selects the reference's own ITensor trainer; not this backend")
        # options that change the reference's results and that the device path does not implement are refused,
        # never silently ignored
        if self.projected_basis and table is None:
            raise ValueError("projected_basis=true: only the projected Legendre bases are real-valued and trainable "
                             "(projected Fourier is complex, loss_functions.jl:343)")
        if self.encode_classes_separately:
            raise ValueError("encode_classes_separately=true (one table per class, Encodings/encodings.jl:57-77) is not "
                             "implemented by the B200 backend; pass precomputed phi (mpst_train_load_phi) instead")
        if np.dtype(self.dtype) != np.dtype(np.float64):
            raise ValueError(f"dtype {self.dtype!r}: the array training path is Float64-only (loss_functions.jl:343)")
        if str(self.svd_alg).lstrip(":") not in ("divide_and_conquer", "qr_iteration", "recursive"):
            raise ValueError(f"unknown svd_alg {self.svd_alg!r}")
        # svd_alg picks the LAPACK driver in the reference; every choice yields the same truncated factors up to
        # rounding, and the device SVD (K5) replaces all of them
        return _ENC[enc] if table is None else "table_" + table[0]


@dataclass
class EncodedTimeSeriesSet:
    """Structs/structs.jl:25-33: here the (class-sorted) scaled series instead of per-sample PStates."""
    X_scaled: np.ndarray            # (T, N) in the encoding range, class-sorted
    original_data: np.ndarray       # (N, T) raw, class-sorted
    labels: np.ndarray              # (N,) sorted labels
    class_distribution: np.ndarray  # counts per sorted class


@dataclass
class TrainedMPS:
    """Structs/options.jl:398-427: (mps, opts, train_data)."""
    mps: list                       # cores, python shapes (chi_l, d, chi_r[, C])
    opts: MPSOptions
    train_data: EncodedTimeSeriesSet
    classes: np.ndarray = None
    enc_table: tuple = None         # (kind, n_sites, ip, dp) of a data-driven encoding: the reference's `encoding_args`


_CTX = {}


def _context(device=None):
    if device is None:
        device = _default_device()
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]


def _default_device():
    import os
    return int(os.environ.get("LOCAL_RANK", "0"))


def _train_opts(opts):
    return make_opts(loss=str(opts.loss_grad).lstrip(":"), bbopt=str(opts.bbopt).lstrip(":"),
                     train_sep=opts.train_classes_separately, update_iters=opts.update_iters, rescale=opts.rescale,
                     chi_max=opts.chi_max, eta=opts.eta, cutoff=opts.cutoff)


def _eval_set(ctx, X_scaled=None, label_idx=None, resident=False):
    """MSE_loss_acc(_conf) (summary.jl:33-114), reduced on the device (mpst_eval_metrics): only 3 + C*C numbers come
    back per set and sweep.  resident=True evaluates the training set already in HBM (on a sharded run each rank
    evaluates its shard and the sums are all-reduced)."""
    if resident:
        sums, conf, n = ctx.eval_metrics()
        rank, world = _dist.rank_world()
        if world > 1:
            import torch
            import torch.distributed as td
            t = torch.tensor(np.concatenate([sums, conf.reshape(-1).astype(np.float64), [float(n)]]),
                             dtype=torch.float64, device=f"cuda:{ctx.device}")
            td.all_reduce(t)
            v = t.cpu().numpy()
            sums, conf, n = v[:3], np.rint(v[3:-1]).astype(np.int64).reshape(conf.shape), int(round(v[-1]))
    else:
        sums, conf, n = ctx.eval_metrics(X_TxN=X_scaled, labels=label_idx)
    return float(sums[0] / n), float(sums[1] / n), float(sums[2] / n), conf


def fitMPS(X_train, y_train=None, X_test=None, y_test=None, opts: Optional[MPSOptions] = None, W=None,
           device=None):
    """fitMPS(X_train, y_train, X_test, y_test, opts) -> (TrainedMPS, info, test_states)
    (RealRealHighDimension.jl:383-562 down to the sweep loop :587-890).  X_*: (N, T) rows are series.
    `W`: optional starting cores (label on the last site), else generate_starting_mps(opts.init_rng)."""
    opts = opts or MPSOptions()
    enc = opts._check()
    X_train = np.asarray(X_train, dtype=np.float64)
    N, T = X_train.shape
    y_train = np.zeros(N, dtype=np.int64) if y_train is None else np.asarray(y_train)
    if not np.issubdtype(y_train.dtype, np.integer):
        raise ValueError("Classes must be integers")                       # :484
    has_test = X_test is not None and np.size(X_test) > 0
    Xs_train, norms = transform_train_data(X_train.T, opts)                # :445
    Xs_sorted, Xo_sorted, ys, _, classes, counts = sort_by_class(Xs_train, X_train, y_train)
    a, b = encoding_range(opts.encoding)
    if not np.all((a <= Xs_sorted) & (Xs_sorted <= b)):
        raise ValueError(f"Data must be rescaled between {a} and {b} before a {opts.encoding} encoding.")
    C = len(classes)
    class_keys = {c: i for i, c in enumerate(classes)}
    train_states = EncodedTimeSeriesSet(Xs_sorted, Xo_sorted, ys, counts)
    test_states = None
    if has_test:
        X_test = np.asarray(X_test, dtype=np.float64)
        y_test = np.asarray(y_test)
        if len(set(np.unique(y_test)) - set(classes)):
            raise ValueError("Test set has classes not present in the training set, this is currently unsupported.")
        Xs_test, _ = transform_test_data(X_test.T, norms, opts)
        Xt_sorted, Xto_sorted, yts, _, _, tcounts = sort_by_class(Xs_test, X_test, y_test)
        test_states = EncodedTimeSeriesSet(Xt_sorted, Xto_sorted, yts, tcounts)

    ctx = _context(device)
    rank, world = _dist.rank_world()
    if world > 1:
        _dist.init_comm(ctx)
        X_local, counts_local, _ = _dist.shard_samples(Xs_sorted, counts, rank, world)
    else:
        X_local, counts_local = Xs_sorted, counts
    # opts.encoding.init on the whole (unsharded) normalised training set (Encodings/encodings.jl:112-120): host work,
    # once per fit; the device evaluates the resulting per-site tables
    enc_table = build_encoding_table(opts, Xs_sorted)
    if enc_table is not None:
        ctx.set_encoding_table(enc_table[0], enc_table[1], opts.d, enc_table[2], enc_table[3])
    ctx.train_load_x(X_local, counts_local, opts.d, opts.chi_max, basis=enc, n_global=N, counts_global=counts)
    cores0 = W if W is not None else generate_starting_mps(opts.chi_init, T, opts.d, C, seed=opts.init_rng)
    ctx.set_cores(cores0)

    info = {"train_loss": [], "train_acc": [], "test_loss": [], "time_taken": [], "train_KL_div": []}
    if has_test:
        info.update({"test_acc": [], "test_KL_div": [], "test_conf": []})
    tr_idx = np.array([class_keys[v] for v in ys])
    te_idx = np.array([class_keys[v] for v in test_states.labels]) if has_test else None

    def log(elapsed):
        if opts.log_level <= 0:
            return None
        mse, kld, acc, _ = _eval_set(ctx, resident=True)
        info["train_loss"].append(mse); info["train_acc"].append(acc)
        info["time_taken"].append(elapsed); info["train_KL_div"].append(kld)
        if has_test:
            mse_t, kld_t, acc_t, conf = _eval_set(ctx, test_states.X_scaled, te_idx)
            info["test_loss"].append(mse_t); info["test_acc"].append(acc_t)
            info["test_KL_div"].append(kld_t); info["test_conf"].append(conf)
        if opts.verbosity > -1:
            print(f"Training KL Div. {kld} | Training acc. {acc}.")
        return acc

    import time
    log(0.0)                                                                # :657-689
    topts = _train_opts(opts)
    ctx.build_env(True)                                                     # :631
    for it in range(opts.nsweeps):                                          # :726
        t0 = time.time()
        for j in range(T - 2, -1, -1):                                      # :731
            ctx.bond_step_quiet(j, True, topts)
        for j in range(T - 1):                                              # :776
            ctx.bond_step_quiet(j, False, topts)
        acc = log(time.time() - t0)                                         # :813-845
        if opts.exit_early and acc == 1.0:                                  # :847-849
            break
    cores = _normalize(ctx.get_cores())                                     # :852
    ctx.set_cores(cores)
    log(float("nan"))                                                       # :854-885
    mps = TrainedMPS(cores, replace(opts), train_states, classes, enc_table)
    return mps, info, test_states


def build_encoding_table(opts, Xs_sorted_TxN):
    """`encoding_args = opts.encoding.init(X_norm, y; opts)` (encodings.jl:112-120) as a device table, or None."""
    spec = opts._table_spec()
    if spec is None:
        return None
    rng_ = encoding_range(opts.encoding)
    if spec[0] == "legendre_proj":
        return _eh.project_legendre(Xs_sorted_TxN, opts.d, norm=spec[1], enc_range=rng_)[0]
    if spec[0] == "sahand_legendre":
        return _eh.init_sahand_legendre(Xs_sorted_TxN, opts.d, time_dependent=spec[1], enc_range=rng_)
    return _eh.split_table(Xs_sorted_TxN, opts.d, opts.aux_basis_dim, aux=spec[2], method=spec[1], enc_range=rng_)


def _normalize(cores):
    """normalize!(W) (RealRealHighDimension.jl:852): unit norm, scale spread over all cores.  Transfer matrices
    E' = sum_s A_s^T E A_s as two BLAS contractions per site (np.einsum's three-operand loop took seconds at chi = 40);
    the running matrix is rescaled at every site so that long chains neither overflow nor underflow."""
    E = np.ones((1, 1, 1))                                     # (a, b, class)
    logn = 0.0
    for A in cores:
        if A.ndim == 4:                                        # label core: (a, s, m, c), E has no class axis yet
            E2 = E[:, :, 0]
            En = np.stack([np.tensordot(A[..., c], np.tensordot(E2, A[..., c], axes=([1], [0])), axes=([0, 1], [0, 1]))
                           for c in range(A.shape[3])], axis=2)
        else:
            En = np.stack([np.tensordot(A, np.tensordot(E[:, :, c], A, axes=([1], [0])), axes=([0, 1], [0, 1]))
                           for c in range(E.shape[2])], axis=2)
        sc = float(np.abs(En).max())
        sc = sc if sc > 0.0 else 1.0
        E = En / sc
        logn += np.log(sc)
    z = np.exp(0.5 * (logn + np.log(float(E[0, 0, :].sum()))) / len(cores))
    return [A / z for A in cores]


def _load_model(ctx, mps: TrainedMPS):
    enc = mps.opts._check()
    T = len(mps.mps)
    C = [A for A in mps.mps if A.ndim == 4][0].shape[3]
    chi = max(max(A.shape[0], A.shape[2]) for A in mps.mps)
    if mps.enc_table is not None:
        ctx.set_encoding_table(mps.enc_table[0], mps.enc_table[1], mps.opts.d, mps.enc_table[2], mps.enc_table[3])
    ctx.model_init(T, C, mps.opts.d, max(chi, 1), basis=enc)
    ctx.set_cores(mps.mps)
    return enc, T, C


def classify(mps: TrainedMPS, X_test, device=None, distributed=False):
    """classify(mps, X_test) -> predicted labels (summary.jl:155-177 -> :116-136): re-derive the train
    normalisation from mps.train_data.original_data, scale the test set with it, argmax_c |yhat_c|^2.
    `distributed`: under torch.distributed every rank classifies its contiguous block of the test series on its own
    GPU (no data-path collective) and the labels are concatenated in rank order on every rank."""
    X_test = np.asarray(X_test, dtype=np.float64)
    if distributed and _dist.rank_world()[1] > 1:
        rank, world = _dist.rank_world()
        b, e = _dist.shard_instances(X_test.shape[0], rank, world)
        local = classify(mps, X_test[b:e], device=device) if e > b else np.asarray(mps.classes)[:0]
        return _dist.gather_instances(np.asarray(local))
    ctx = _context(device)
    _load_model(ctx, mps)
    _, norms = transform_train_data(mps.train_data.original_data.T, mps.opts)     # :159-160
    Xs, _ = transform_test_data(X_test.T, norms, mps.opts)
    _, am = ctx.overlaps(X_TxN=Xs)
    return np.asarray(mps.classes)[am]


# ------------------------------------------------------------------------------------------------
# imputation (Imputation/imputation.jl)
# ------------------------------------------------------------------------------------------------
@dataclass
class ImputationProblem:
    """imputation.jl:10-20."""
    mps: TrainedMPS
    X_train: np.ndarray
    y_train: np.ndarray
    X_test: np.ndarray
    y_test: np.ndarray
    opts: MPSOptions
    xvals: np.ndarray
    class_map: dict
    norms: object = None
    train_mean: float = 0.0


def make_grid(enc_range, dx):
    """collect(range(a, b; step=dx)) (imputation.jl:90), correctly rounded like Julia's ranges."""
    a, b = enc_range
    G = int(np.floor((b - a) / dx + 1e-9)) + 1
    inv = round(1.0 / dx)
    if abs(inv * dx - 1.0) < 1e-12 and float(a * inv).is_integer():
        return (a * inv + np.arange(G)) / inv
    return a + np.arange(G) * dx


def init_imputation_problem(mps: TrainedMPS, X_test, y_test=None, dx=1e-4, guess_range=None, verbosity=1):
    """init_imputation_problem(mps, X_test, y_test) (imputation.jl:143-190 -> 48-123)."""
    opts = mps.opts
    enc = opts._check()
    X_test = np.asarray(X_test, dtype=np.float64)
    y_test = np.zeros(X_test.shape[0], dtype=np.int64) if y_test is None else np.asarray(y_test)
    X_train = mps.train_data.original_data
    y_train = mps.train_data.labels
    rng_ = guess_range or encoding_range(opts.encoding)
    xvals = make_grid(rng_, dx)
    _, norms = transform_train_data(X_train, opts)          # hoisted out of get_predictions (:287)
    class_map = {c: i for i, c in enumerate(sorted(np.unique(y_train)))}
    return ImputationProblem(mps, X_train, y_train, X_test, y_test, opts, xvals, class_map, norms,
                             float(np.mean(X_train)))


def get_predictions_batch(imp: ImputationProblem, cls, instances, missing_sites_list, method="median",
                          invert_transform=True, rseed=1, num_trajectories=1, max_jump=None, uniforms=None,
                          device=None, impute_order="forwards", get_wmad=False, get_std=False,
                          rejection_threshold=None, max_trials=10, return_err=False, distributed=False):
    """Batched get_predictions (imputation.jl:264-410): instances are indices into the test series of
    class `cls`; missing_sites_list[k] are the 0-based sites to impute in instance k.  Keyword arguments are those of
    impute_median / impute_mean / impute_mode / impute_ITS (MPS_methods.jl:201-347); `rejection_threshold=None` is the
    reference's `:none`.  Returns (ts (n, n_traj, T), target (n, T)), or (ts, pred_err, target) with return_err:
    pred_err (n, n_traj, T) holds WMAD (median, get_wmad) / std (mean, get_std) at the imputed sites, transformed back
    like the reference does (:337-384: add the series, invert, subtract; values the logit cannot invert become NaN)."""
    if method not in ("median", "mean", "mode", "ITS"):
        raise ValueError("Invalid method. Choose :mean, :mode, :median or :ITS")
    if distributed and _dist.rank_world()[1] > 1:
        # one process per GPU: every rank imputes its own contiguous block of instances (no communication on the data
        # path), then the results are concatenated in rank order on every rank
        rank, world = _dist.rank_world()
        b, e = _dist.shard_instances(len(instances), rank, world)
        if uniforms is None and method == "ITS":
            # draw the whole batch's stream on every rank so that the result does not depend on the number of ranks
            Kmax_all = max((len(ms) for ms in missing_sites_list), default=0)
            per_site = max_trials if rejection_threshold is not None else 1
            uniforms = np.random.RandomState(rseed).random_sample((len(instances), num_trajectories, Kmax_all * per_site))
        local = get_predictions_batch(imp, cls, list(instances)[b:e], list(missing_sites_list)[b:e], method=method,
                                      invert_transform=invert_transform, rseed=rseed, num_trajectories=num_trajectories,
                                      max_jump=max_jump, uniforms=None if uniforms is None else np.asarray(uniforms)[b:e],
                                      device=device, impute_order=impute_order, get_wmad=get_wmad, get_std=get_std,
                                      rejection_threshold=rejection_threshold, max_trials=max_trials, return_err=return_err)
        return _dist.gather_instances(tuple(local))
    ctx = _context(device)
    _load_model(ctx, imp.mps)
    opts = imp.opts
    cl_inds = np.nonzero(imp.y_test == cls)[0]
    n = len(instances)
    T = imp.X_test.shape[1]
    if n == 0:                                                             # an empty block of a sharded batch
        nt = num_trajectories if method == "ITS" else 1
        z = np.zeros((0, nt, T))
        return (z, z.copy(), np.zeros((0, T))) if return_err else (z, np.zeros((0, T)))
    raw = imp.X_test[cl_inds[np.asarray(instances, dtype=np.int64)]]       # (n, T)
    filled = raw.copy()
    mask = np.zeros((n, T), dtype=np.uint8)
    for k, ms in enumerate(missing_sites_list):
        filled[k, list(ms)] = imp.train_mean                               # :290
        mask[k, list(ms)] = 1
    Xs, oob = transform_test_data(filled.T, imp.norms, opts)               # :291
    Kmax = int(mask.sum(axis=1).max()) if n else 0
    if method == "ITS" and uniforms is None:
        # the reference draws rand(MersenneTwister(rseed)) site by site (MPS_methods.jl:324); callers
        # needing Julia's stream pass `uniforms` drawn in Julia
        rs = np.random.RandomState(rseed)
        per_site = max_trials if rejection_threshold is not None else 1
        uniforms = rs.random_sample((n, num_trajectories, Kmax * per_site))
    if method == "ITS" and rejection_threshold is None:
        # plain ITS reads uniforms[i, trajectory, k] for the k-th missing site: a wider array (drawn for a larger batch) is cut
        uniforms = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(n, num_trajectories, -1)[:, :, :Kmax])
    get_err = (method == "median" and get_wmad) or (method == "mean" and get_std)
    out, err = ctx.impute_batch(imp.class_map[cls], Xs, mask.T, imp.xvals, method=method, uniforms=uniforms,
                                n_traj=num_trajectories if method == "ITS" else 1,
                                max_jump=-1.0 if max_jump is None else float(max_jump), impute_order=impute_order,
                                get_err=get_err, rejection_threshold=rejection_threshold if method == "ITS" else None,
                                max_trials=max_trials, return_err=True)
    if invert_transform:                                                   # :337-394
        res = np.empty_like(out)
        for tr in range(out.shape[1]):
            res[:, tr, :] = invert_test_transform(out[:, tr, :].T, oob, imp.norms, opts).T
            if method in ("median", "mean"):                               # :340-379
                with np.errstate(invalid="ignore"):
                    err[:, tr, :] = invert_test_transform((err[:, tr, :] + out[:, tr, :]).T, oob, imp.norms, opts).T - res[:, tr, :]
        return (res, err, raw) if return_err else (res, raw)
    full, _ = transform_test_data(raw.T, imp.norms, opts)
    return (out, err, full.T) if return_err else (out, full.T)


def MPS_impute(imp: ImputationProblem, cls, instance, missing_sites, method="median", **kw):
    """MPS_impute(imp, class, instance, missing_sites, method) (imputation.jl:467-563) without the
    plotting / kNN-baseline extras: returns (imputed_ts list, pred_err, target, stats)."""
    ts, err, target = get_predictions_batch(imp, cls, [instance], [missing_sites], method=method, return_err=True, **kw)
    ts = [ts[0, t] for t in range(ts.shape[1])]
    ms = list(missing_sites)
    stats = [{"MAE": float(np.mean(np.abs(t[ms] - target[0][ms]))),
              "MAPE": float(np.mean(np.abs((t[ms] - target[0][ms]) / target[0][ms])))} for t in ts]
    # imputation.jl:299-318: only :mean and :median produce error bars; the others return `nothing` per trajectory
    pred_err = [err[0, t] for t in range(len(ts))] if method in ("median", "mean") else [None for _ in ts]
    return ts, pred_err, target[0], stats


class MPSClassifier:
    """MLJ `MPSClassifier` (MLJIntegration/MLJ_integration.jl:2-62): same hyper-parameter fields and
    defaults; fit/predict route to fitMPS/classify."""

    def __init__(self, nsweeps=5, chi_max=15, eta=0.01, d=2, encoding="Legendre_No_Norm", projected_basis=False,
                 aux_basis_dim=2, cutoff=1e-10, update_iters=1, loss_grad="KLD", bbopt="TSGO", rescale=(False, True),
                 train_classes_separately=False, encode_classes_separately=False, minmax=True, exit_early=True,
                 sigmoid_transform=True, init_rng=1234, chi_init=4, reformat_verbosity=-1):
        self.__dict__.update({k: v for k, v in locals().items() if k != "self"})
        self.fitresult = None

    def _opts(self):
        f = dict(self.__dict__)
        f.pop("fitresult")
        f.pop("reformat_verbosity")
        return MPSOptions(verbosity=-1, log_level=0, **f)

    def fit(self, X, y):
        self.fitresult, _, _ = fitMPS(X, y, opts=self._opts())
        return self

    def predict(self, X):
        return classify(self.fitresult, X)
