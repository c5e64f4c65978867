"""CPU Float64 oracle for the MPSTime hot path (fitMPS sweep, classify, MPS_impute).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may use it, and there only as the checker / the timed CPU baseline.

It is a plain-numpy *restatement* of the reference's algorithm (the reference is Julia and
cannot run in this image, so nothing is compiled from /root/reference).  Every function cites
the reference file:line it follows (paths relative to /root/reference/src).

Parity status
-------------
PINNED against the reference's own serialized output (`test/Data/ecg200/mps_saves/test_dataset.jld2`, the only
artefact under /root/reference that holds numbers of this path; extracted by `tests/golden/make_golden_from_jld2.py`
-> `ecg200_legendre.npz` and `tests/golden/make_golden_mps_from_jld2.py` -> `ecg200_trained_mps.npz`):
* the RobustSigmoid -> MinMax normalisation and the Legendre encoding values (max deviation 2e-15);
* the core index order (left link, site, right link[, class]), the label position, `contract_mps` / `classify`
  (the reference's trained MPS classifies the reference's own encoded training set 100/100 through `overlaps`),
  `normalize!` (<W|W> = 1 to 1e-15) and the canonical form a sweep leaves behind: sites 1..T-1 left-orthonormal (U of
  `decomposeBT`), the label core's Gram matrix diagonal with descending entries (V*S, LAPACK ordering)
  (tests/test_oracle.py::test_reference_trained_mps_pins_layout_norm_and_classification).
PARITY UNPINNED (no replayable reference output exists: `test/classification.jl:26,47`, `test/imputation.jl:34-52` need
the UCR downloads and a missing blob; the arithmetic lives in un-vendored Julia packages whose published algorithms
are restated -- ITensors 0.6.22 / NDTensors 0.3.74 `svd` + `truncate!`, NumericalIntegration 0.2.0 `cumul_integrate`,
Normalization 0.7.3, StatsBase 0.34.4 weighted `median`, KernelDensity / Interpolations for the data-driven bases):
the per-bond loss / gradient values, the TSGO / GD update, the truncation rule's cutoff decision, the environment
update, and the imputation step (median / mean / mode / ITS, WMAD, rejection, backwards order).  These follow the
reference source line by line and are self-checked (literal loop == vectorised form, gradient == finite differences,
backwards == forwards on the mirrored chain, C restatement == numpy).

Conventions (0-based sites j = 0..T-1)
--------------------------------------
cores[j]  : ndarray (chi_j, d, chi_{j+1}) with chi_0 = chi_T = 1; the label core carries a
            trailing class axis: (chi_j, d, chi_{j+1}, C).
phi       : (N, T, d) encoded samples, sorted by class; counts[c] = samples in class c.
B         : bond tensor as the reference's `BondTensor = Matrix` (D x C); column c is the
            flattening with s_l fastest: idx = s_l + d*(a + chi_l*(s_r + d*b))
            (loss_functions.jl:193-200, 248-262, 294).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# A.1  normalisation  (utils.jl:161-200, 202-278, 299-330; options.jl:72-84)
# --------------------------------------------------------------------------------------

ENCODING_RANGE = {
    "legendre": (-1.0, 1.0), "legendre_no_norm": (-1.0, 1.0), "legendre_norm": (-1.0, 1.0),
    "fourier": (-1.0, 1.0), "stoudenmire": (0.0, 1.0), "sahand": (0.0, 1.0), "uniform": (0.0, 1.0),
}


def fit_train_norms(X, sigmoid_transform=True, minmax=True):
    """transform_train_data (utils.jl:161-200): global RobustSigmoid then global MinMax.
    Normalization.jl RobustSigmoid params = (median, IQR) with Julia type-7 quantiles
    (numpy 'linear'); formula options.jl:72-77.  Returns (Xs in [0,1]-ish, norms dict)."""
    X = np.asarray(X, dtype=np.float64)
    norms = {"sigmoid": None, "minmax": None}
    Xs = X.copy()
    if sigmoid_transform:
        m = float(np.median(X))
        q75, q25 = np.quantile(X, [0.75, 0.25])
        iqr = float(q75 - q25)
        norms["sigmoid"] = (m, iqr)
        Xs = 1.0 / (1.0 + np.exp(-(Xs - m) / (iqr / 1.35)))
    if minmax:
        lo, hi = float(Xs.min()), float(Xs.max())
        norms["minmax"] = (lo, hi)
        Xs = (Xs - lo) / (hi - lo)
    return Xs, norms


def transform_train_data(X, enc_range=(-1.0, 1.0), sigmoid_transform=True, minmax=True,
                         data_bounds=(0.0, 1.0)):
    """utils.jl:161-200.  X holds series as columns (T x N) but every statistic is global."""
    Xs, norms = fit_train_norms(X, sigmoid_transform, minmax)
    if minmax:
        lb, ub = data_bounds
        Xs = Xs * (ub - lb) + lb                      # utils.jl:181-192
    a, b = enc_range
    Xs = (b - a) * Xs + a                             # utils.jl:196-197
    return Xs, norms


def transform_test_data(X, norms, enc_range=(-1.0, 1.0), minmax=True, data_bounds=(0.0, 1.0),
                        rescale_out_of_bounds=True):
    """utils.jl:202-278.  X is (T x n) (series are columns) or a single series (T,).
    Returns (X_scaled, oob_rescales) with oob_rescales = [(col, lb_shift, ub_scale), ...]."""
    X = np.asarray(X, dtype=np.float64)
    single = X.ndim == 1
    Xs = X.reshape(-1, 1).copy() if single else X.copy()
    if Xs.size == 0:
        return Xs, []
    if norms["sigmoid"] is not None:                  # utils.jl:221-223
        m, iqr = norms["sigmoid"]
        Xs = 1.0 / (1.0 + np.exp(-(Xs - m) / (iqr / 1.35)))
    if norms["minmax"] is not None:
        lo, hi = norms["minmax"]
        Xs = (Xs - lo) / (hi - lo)
    if minmax:                                        # utils.jl:226-235
        lb, ub = data_bounds
        Xs = Xs * (ub - lb) + lb
    oob = []
    if rescale_out_of_bounds:                         # utils.jl:243-265
        for i in range(Xs.shape[1]):
            col = Xs[:, i]
            lb_s, ub_s = 0.0, 1.0
            lo, hi = col.min(), col.max()
            if lo < 0:
                col -= lo
                hi = col.max()
                lb_s = float(lo)
            if hi > 1:
                col /= hi
                ub_s = float(hi)
            if (lb_s, ub_s) != (0.0, 1.0):
                oob.append((i, lb_s, ub_s))
    a, b = enc_range
    Xs = (b - a) * Xs + a                             # utils.jl:274-275
    return (Xs[:, 0] if single else Xs), oob


def invert_test_transform(Xs, oob, norms, enc_range=(-1.0, 1.0), minmax=True,
                          data_bounds=(0.0, 1.0)):
    """utils.jl:299-330 (inverse sigmoid = logit; NaN where the argument leaves (0,1), the
    reference raises DomainError there and NaNs the entry, imputation.jl:344-384)."""
    Xs = np.asarray(Xs, dtype=np.float64)
    single = Xs.ndim == 1
    X = Xs.reshape(-1, 1).copy() if single else Xs.copy()
    a, b = enc_range
    X = (X - a) / (b - a)
    for (i, lb_s, ub_s) in oob:
        X[:, i] = X[:, i] * ub_s + lb_s
    if minmax:
        lb, ub = data_bounds
        X = (X - lb) / (ub - lb)
    if norms["minmax"] is not None:
        lo, hi = norms["minmax"]
        X = X * (hi - lo) + lo
    if norms["sigmoid"] is not None:
        m, iqr = norms["sigmoid"]
        with np.errstate(invalid="ignore", divide="ignore"):
            X = np.where((X > 0) & (X < 1), np.log(X / (1.0 - X)), np.nan) * (iqr / 1.35) + m
    return X[:, 0] if single else X


# --------------------------------------------------------------------------------------
# A.2  encodings  (Encodings/bases.jl:13-92)
# --------------------------------------------------------------------------------------

def legendre_encode(x, d, norm=False):
    """bases.jl:77-92: phi_l(x) = sqrt((2l+1)/2) P_l(x), l = 0..d-1 (LegendrePolynomials.jl
    `Pl(x,l; norm=Val(:normalized))`, Bonnet recurrence).  norm=True divides by
    sqrt(Pl(1,d;normalized)*d) = sqrt(sqrt((2d+1)/2)*d)  (:86-89)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty(x.shape + (d,), dtype=np.float64)
    p_prev = np.ones_like(x)
    out[..., 0] = np.sqrt(0.5) * p_prev
    if d > 1:
        p = x.copy()
        out[..., 1] = np.sqrt(1.5) * p
        for l in range(1, d - 1):
            p_next = ((2 * l + 1) * x * p - l * p_prev) / (l + 1)
            p_prev, p = p, p_next
            out[..., l + 1] = np.sqrt((2 * (l + 1) + 1) / 2.0) * p
    if norm:
        out /= np.sqrt(np.sqrt((2 * d + 1) / 2.0) * d)
    return out


def fourier_freqs(d):
    """get_fourier_freqs bases.jl:27-34: [0, 1, -1, 2, -2, ...][:d]."""
    hb = int(np.ceil((d - 1.0) / 2.0))
    f = [0]
    for i in range(1, hb + 1):
        f += [i, -i]
    return np.array(f[:d], dtype=np.float64)


def fourier_encode(x, d):
    """bases.jl:23-42: cispi(f_k x)/sqrt(d) (complex)."""
    x = np.asarray(x, dtype=np.float64)
    return np.exp(1j * np.pi * x[..., None] * fourier_freqs(d)) / np.sqrt(d)


def stoudenmire_encode(x):
    """angle_encode bases.jl:13-20 (d=2, x in [0,1], periods=1/4)."""
    x = np.asarray(x, dtype=np.float64)
    s1 = np.exp(1j * np.pi * 1.5 * x) * np.cos(np.pi * 0.5 * x)
    s2 = np.exp(-1j * np.pi * 1.5 * x) * np.sin(np.pi * 0.5 * x)
    return np.stack([s1, s2], axis=-1)


def sahand_encode(x, d):
    """bases.jl:53-74 (d even, x in [0,1])."""
    assert d % 2 == 0
    x = np.asarray(x, dtype=np.float64)
    dx = 2.0 / d
    out = np.zeros(x.shape + (d,), dtype=np.complex128)
    for i in range(1, d + 1):
        interval = np.ceil(i / 2.0)
        startx = (interval - 1) * dx
        inside = (startx <= x) & (x <= interval * dx)
        if i % 2 == 1:
            s = np.exp(1j * np.pi * 1.5 * x / dx) * np.cos(np.pi * 0.5 * (x - startx) / dx)
        else:
            s = np.exp(-1j * np.pi * 1.5 * x / dx) * np.sin(np.pi * 0.5 * (x - startx) / dx)
        out[..., i - 1] = np.where(inside, s, 0.0)
    return out


def uniform_encode(x, d):
    """bases.jl:2-5."""
    x = np.asarray(x)
    return np.full(x.shape + (d,), 1.0 / d)


def encode(x, d, basis="legendre_no_norm"):
    b = basis.lower()
    if b in ("legendre", "legendre_no_norm"):
        return legendre_encode(x, d, norm=False)
    if b == "legendre_norm":
        return legendre_encode(x, d, norm=True)
    if b == "fourier":
        return fourier_encode(x, d)
    if b == "stoudenmire":
        assert d == 2
        return stoudenmire_encode(x)
    if b == "sahand":
        return sahand_encode(x, d)
    if b == "uniform":
        return uniform_encode(x, d)
    raise ValueError(basis)


# --------------------------------------------------------------------------------------
# data-driven / time-dependent encodings GIVEN their initialised arguments (bases.jl:95-129,
# splitbases.jl:96-163).  The initialisers themselves (KDE, series projections, histogram bins) are
# host-side in the reference and in the product; the oracle restates what is evaluated per point.
# --------------------------------------------------------------------------------------

def projected_legendre_encode(x, orders, norm=False):
    """legendre_encode(x, nds, ds; norm) (bases.jl:95-108): the normalised Legendre polynomials of the given
    orders; with norm the vector is divided by sqrt(Pl(1, dmax) * dmax), dmax = maximum(ds)."""
    orders = np.asarray(orders, dtype=np.int64)
    full = legendre_encode(np.asarray(x, dtype=np.float64), int(orders.max()) + 1)
    out = full[..., orders]
    if norm:
        dmax = int(orders.max())
        out = out / np.sqrt(np.sqrt((2 * dmax + 1) / 2.0) * dmax)
    return out


def interp_kde_pdf(x, kde_x, coeffs):
    """pdf(kde, x) (KernelDensity.InterpKDE, un-vendored): quadratic B-spline through the tabulated density
    (Interpolations BSpline(Quadratic(Line(OnGrid()))), padded coefficients), zero outside the grid."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    out = np.zeros_like(x)
    n = len(kde_x)
    h = kde_x[1] - kde_x[0]
    for k, xv in enumerate(x):
        t = (xv - kde_x[0]) / h + 1.0
        if t < 1.0 or t > n:
            continue
        i = int(np.rint(t))
        dx = t - i
        out[k] = coeffs[i - 1] * (dx - 0.5) ** 2 / 2 + coeffs[i] * (0.75 - dx ** 2) + coeffs[i + 1] * (dx + 0.5) ** 2 / 2
    return out


def sahand_legendre_encode(x, d, kde_x, coeffs, minx, scale, cVecs):
    """sahand_legendre_encode (bases.jl:111-117): f_n(x) = (sum_i cVecs[n,i] x^(i-1)) * f0(x) / scale,
    f0 = max(sqrt(max(pdf(kde, x), 0)), minx)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    f0 = np.maximum(np.sqrt(np.maximum(interp_kde_pdf(x, kde_x, coeffs), 0.0)), minx)
    powers = x[:, None] ** np.arange(d)[None, :]
    return (powers @ np.asarray(cVecs).T) * f0[:, None] / scale


def split_encode(x, bins, aux_dim, aux_basis="uniform"):
    """project_onto_bins (splitbases.jl:113-140) with rect (:96-109): the auxiliary basis, evaluated at the position
    inside the bin stretched to the whole range, in the slot of the bin containing x; a point exactly on an inner
    edge gets weight 1/2 in both neighbours."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    bins = np.asarray(bins, dtype=np.float64)
    nb = len(bins) - 1
    a, b = bins[0], bins[-1]
    scale = b - a
    out = np.zeros((len(x), nb * aux_dim))
    for k, xv in enumerate(x):
        for i in range(nb):
            dxb = bins[i + 1] - bins[i]
            xprop = scale * (xv - bins[i]) / dxb
            r = xprop / scale - 0.5
            lb = 1.0 if i == 0 else 0.5
            rb = 1.0 if i == nb - 1 else 0.5
            sel = lb if r == -0.5 else rb if r == 0.5 else 1.0 if -0.5 <= r <= 0.5 else 0.0
            if sel != 0.0:
                out[k, i * aux_dim:(i + 1) * aux_dim] = sel * encode(np.array([a + xprop]), aux_dim, aux_basis)[0]
    return out


def encode_dataset(X_scaled, y, d, basis="legendre_no_norm"):
    """encode_dataset -> encode_safe_dataset (encodings.jl:33-46, 79-156): stable sort by class
    (`sortperm`), encode every (sample, site).  X_scaled is T x N (series are columns).
    Returns (phi (N,T,d), y_sorted, order, counts, classes)."""
    X_scaled = np.asarray(X_scaled, dtype=np.float64)
    y = np.asarray(y)
    order = np.argsort(y, kind="stable")
    Xs = X_scaled[:, order]
    ys = y[order]
    a, b = ENCODING_RANGE[basis.lower()]
    if not np.all((a <= Xs) & (Xs <= b)):                     # encodings.jl:117-121
        raise ValueError(f"Data must be rescaled between {a} and {b} before encoding")
    phi = encode(Xs.T, d, basis)                              # (N, T, d)
    classes, counts = np.unique(ys, return_counts=True)       # encodings.jl:151-152
    return phi, ys, order, counts.astype(np.int64), classes


# --------------------------------------------------------------------------------------
# starting MPS (RealRealHighDimension.jl:1-41) -- NOT bit-compatible with ITensors' random_mps
# (un-vendored, Julia RNG): the reference-side shim keeps generate_startingMPS in Julia and
# passes the cores in.  This generator only needs the same *structure*.
# --------------------------------------------------------------------------------------

def _norm2_general(cores):
    E = np.ones((1, 1, 1))            # (a, b, c) with c the (possibly trivial) class axis
    for A in cores:
        if A.ndim == 4:
            assert E.shape[2] == 1
            E = np.einsum("ab,asmc,bsnc->mnc", E[:, :, 0], A, A)
        else:
            E = np.einsum("abc,asm,bsn->mnc", E, A, A)
    return float(E[0, 0, :].sum())


def normalize_mps(cores):
    """ITensors `normalize!(W)` (RealRealHighDimension.jl:32,852): rescale to unit norm, the
    scale spread evenly over all cores (z = norm^(1/T)).  Un-vendored; predictions and
    imputations are invariant to how the scale is distributed."""
    n2 = _norm2_general(cores)
    z = np.exp(0.5 * np.log(n2) / len(cores))
    return [A / z for A in cores]


def random_start_mps(T, d, chi_init, C, seed=1234):
    """Structure of generate_startingMPS (RealRealHighDimension.jl:1-41): random MPS with
    uniform link dimension chi_init (capped by what the sites can support), class index of
    dimension C attached to the LAST site, normalised, orthogonality centre at the last site
    (all other cores left-orthonormal)."""
    rng = np.random.default_rng(seed)
    chis = [1]
    for j in range(1, T):
        chis.append(int(min(chi_init, d ** min(j, T - j), 2 ** 62)))
    chis.append(1)
    cores = []
    for j in range(T):
        shape = (chis[j], d, chis[j + 1]) + ((C,) if j == T - 1 else ())
        cores.append(rng.standard_normal(shape))
    # orthogonalize!(W, T): QR sweep left -> right  (:37)
    for j in range(T - 1):
        A = cores[j]
        a, s, b = A.shape
        Q, R = np.linalg.qr(A.reshape(a * s, b))
        k = Q.shape[1]
        cores[j] = Q.reshape(a, s, k)
        nxt = np.tensordot(R, cores[j + 1], axes=(1, 0))
        cores[j + 1] = nxt / np.linalg.norm(nxt)             # keep the running scale at 1 (long chains overflow otherwise)
    n2 = _norm2_general(cores)
    cores[-1] = cores[-1] / np.sqrt(n2)
    return cores


# --------------------------------------------------------------------------------------
# environments  (RealRealHighDimension.jl:45-103 construct_caches, 107-144 update_caches!)
# --------------------------------------------------------------------------------------

def env_step_left(phi_j, LE_prev, core):
    """LE[j,i][k] = sum_{s,a} x[s] core[a,s,k] LE[j-1,i][a]   (:75, :139; first site :69,:137)."""
    return np.einsum("is,ia,ask->ik", phi_j, LE_prev, core, optimize=True)


def env_step_right(phi_j, RE_next, core):
    """RE[j,i][k] = sum_{s,b} x[s] core[k,s,b] RE[j+1,i][b]   (:93, :129; last site :84,:127)."""
    return np.einsum("is,ib,ksb->ik", phi_j, RE_next, core, optimize=True)


def construct_caches(cores, phi, going_left=True):
    """construct_caches (:45-103).  going_left=True builds LE[0..T-2]; False builds RE[T-1..1].
    Returns dict site -> (N, chi) array."""
    N, T, d = phi.shape
    ones = np.ones((N, 1))
    env = {}
    if going_left:
        prev = ones
        for j in range(T - 1):
            prev = env_step_left(phi[:, j], prev, cores[j])
            env[j] = prev
    else:
        nxt = ones
        for j in range(T - 1, 0, -1):
            nxt = env_step_right(phi[:, j], nxt, cores[j])
            env[j] = nxt
    return env


# --------------------------------------------------------------------------------------
# bond tensor flatten / loss+grad / update / split
# --------------------------------------------------------------------------------------

def flatten_bt(core_l, core_r):
    """flatten_bt (RealRealHighDimension.jl:221-238): B[:,c] = vec(W[l]*W[r]*onehot(c)) in the
    layout the loss kernels index (s_l fastest).  Returns (B (D,C), (chi_l, d, chi_r))."""
    if core_l.ndim == 4:
        full = np.einsum("asmc,mtb->btasc", core_l, core_r)
    else:
        full = np.einsum("asm,mtbc->btasc", core_l, core_r)
    chi_r, d, chi_l, _, C = full.shape
    return full.reshape(chi_r * d * chi_l * d, C), (chi_l, d, chi_r)


def phi_tilde(xl, L, xr, R):
    """phi~ = kron layout of kron_scaleadd_KLD! (loss_functions.jl:248-262): xa = kron(R, xr)
    (second arg fastest), xb = kron(L, xl); idx = j_b + len(xb)*i_a."""
    xa = np.kron(R, xr)
    xb = np.kron(L, xl)
    return np.kron(xa, xb)


def loss_grad_KLD_loop(B, L, R, xl, xr, counts, train_sep=False):
    """Literal, sample-sequential restatement of Loss_Grad_KLD (loss_functions.jl:322-432)
    including the software-pipelined accumulation k += kprev/scale; kprev = phi (:203-215)
    and the final flush (:367).  Small N only."""
    D, C = B.shape
    N = xl.shape[0]
    grad = np.zeros((D, C))
    losses = 0.0
    i0 = 0
    for ci, cn in enumerate(counts):
        yhat = 1.0
        kprev = np.zeros(D)
        k = np.zeros(D)
        loss = 0.0
        for i in range(i0, i0 + cn):
            ph = phi_tilde(xl[i], L[i], xr[i], R[i])
            k += kprev / yhat
            kprev = ph
            yhat = float(B[:, ci] @ ph)
            loss += -np.log(yhat * yhat)                      # KLD_iter! :318
        if train_sep:                                         # :424-425
            losses += loss / cn
            grad[:, ci] = -(k + kprev / yhat) / cn
        else:                                                 # :366-367
            losses += loss
            grad[:, ci] = -(k + kprev / yhat) / N
        i0 += cn
    if not train_sep:
        losses /= N                                           # :371
    return losses, grad


def _PQ(xl, L, xr, R):
    n = xl.shape[0]
    P = (L[:, :, None] * xl[:, None, :]).reshape(n, -1)       # p = s_l + d*a
    Q = (R[:, :, None] * xr[:, None, :]).reshape(n, -1)       # q = s_r + d*b
    return P, Q


def bond_yhat(B, L, R, xl, xr):
    """yhat[i,c] = <B_c, phi~_i> for every sample and class (loss_functions.jl:210,256)."""
    P, Q = _PQ(xl, L, xr, R)
    D, C = B.shape
    out = np.empty((P.shape[0], C))
    for c in range(C):
        Bc = B[:, c].reshape(Q.shape[1], P.shape[1])
        out[:, c] = np.einsum("iq,qp,ip->i", Q, Bc, P, optimize=True)
    return out


def loss_grad_KLD(B, L, R, xl, xr, counts, train_sep=False):
    """Vectorised Loss_Grad_KLD (loss_functions.jl:322-432): each class column sees only its own
    samples; loss = (1/N) sum -log(yhat^2), G_c = -(1/N) sum_{i in I_c} phi~_i / yhat_i."""
    D, C = B.shape
    N = xl.shape[0]
    P, Q = _PQ(xl, L, xr, R)
    grad = np.zeros((D, C))
    loss = 0.0
    i0 = 0
    for c, cn in enumerate(counts):
        sl = slice(i0, i0 + cn)
        Bc = B[:, c].reshape(Q.shape[1], P.shape[1])
        yh = np.einsum("iq,qp,ip->i", Q[sl], Bc, P[sl], optimize=True)
        lc = float(np.sum(-np.log(yh * yh)))
        denom = cn if train_sep else N
        loss += lc / denom
        grad[:, c] = (-(Q[sl] / yh[:, None]).T @ P[sl]).reshape(-1) / denom
        i0 += cn
    return loss, grad


def loss_grad_MSE_loop(B, L, R, xl, xr, counts):
    """Literal restatement of Loss_Grad_MSE (loss_functions.jl:561-619) with the pipelined
    accumulation k += kprev*(yhat - yprev) (:435-496)."""
    D, C = B.shape
    N = xl.shape[0]
    grad = np.zeros((D, C))
    losses = 0.0
    mask = np.zeros(N)
    yprev = 0.0
    i0 = 0
    for ci, cn in enumerate(counts):
        yhat = 1.0
        kprev = np.zeros(D)
        k = np.zeros(D)
        mask[i0:i0 + cn] = 1.0
        loss = 0.0
        for i in range(N):
            ph = phi_tilde(xl[i], L[i], xr[i], R[i])
            k += kprev * (yhat - yprev)
            kprev = ph
            yhat = float(B[:, ci] @ ph)
            loss += 0.5 * (yhat - mask[i]) ** 2               # MSE_iter! :553
            yprev = mask[i]
        losses += loss
        grad[:, ci] = (k + kprev * (yhat - yprev)) / N        # :610
        mask[i0:i0 + cn] = 0.0
        i0 += cn
    return losses / N, grad


def loss_grad_MSE(B, L, R, xl, xr, counts):
    """Vectorised Loss_Grad_MSE (loss_functions.jl:561-619): every class column sees all samples."""
    D, C = B.shape
    N = xl.shape[0]
    P, Q = _PQ(xl, L, xr, R)
    yh = bond_yhat(B, L, R, xl, xr)
    onehot = np.zeros((N, C))
    i0 = 0
    for c, cn in enumerate(counts):
        onehot[i0:i0 + cn, c] = 1.0
        i0 += cn
    diff = yh - onehot
    loss = float(0.5 * np.sum(diff * diff)) / N
    grad = np.empty((D, C))
    for c in range(C):
        grad[:, c] = ((Q * diff[:, c:c + 1]).T @ P).reshape(-1) / N
    return loss, grad


def apply_update(B, L, R, xl, xr, counts, loss="KLD", bbopt="TSGO", eta=0.01, update_iters=1,
                 rescale=(False, True), train_sep=False, loop=False):
    """apply_update / TSGO / custGD (loss_functions.jl:27-188).  Returns (B_new, loss_first,
    gradnorm_first) where loss/gradnorm are those of the first iteration (what fitMPS prints
    with track_cost)."""
    B = B.copy()
    if rescale[0]:
        B /= np.linalg.norm(B)                                # :109-111
    l0 = g0 = None
    for _ in range(update_iters):
        if loss.upper() == "KLD":
            fn = loss_grad_KLD_loop if loop else loss_grad_KLD
            lo, G = fn(B, L, R, xl, xr, counts, train_sep)
        else:
            fn = loss_grad_MSE_loop if loop else loss_grad_MSE
            lo, G = fn(B, L, R, xl, xr, counts)
        gn = float(np.linalg.norm(G))
        if l0 is None:
            l0, g0 = lo, gn
        if bbopt.upper() == "TSGO":
            B -= eta * G / gn                                 # :79
        else:
            B -= eta * G                                      # :49
    if rescale[1]:
        B /= np.linalg.norm(B)                                # :177-179
    return B, l0, g0


def truncate_spectrum(P, maxdim, cutoff, mindim=1):
    """NDTensors 0.3.74 `truncate!` (un-vendored; published algorithm restated): P = sigma^2
    descending; drop from the tail while n > maxdim, then while the *sum* of discarded weight
    (including what maxdim already discarded) + P[n] <= cutoff*sum(P) and n > mindim."""
    P = np.asarray(P, dtype=np.float64).copy()
    P[P < 0] = 0.0
    n = len(P)
    if n == 1:
        return 1
    err = 0.0
    while n > maxdim:
        err += P[n - 1]
        n -= 1
    scale = float(P.sum())
    if scale == 0.0:
        scale = 1.0
    while n > mindim and err + P[n - 1] <= cutoff * scale:
        err += P[n - 1]
        n -= 1
    return max(n, 1)


def decompose_bt(B, dims, going_left, chi_max, cutoff):
    """decomposeBT (RealRealHighDimension.jl:146-203) -> ITensors.svd (LAPACK gesdd) + truncate.
    going_left : SVD of M[(a,c,s_l),(s_r,b)], W[l] <- U*S (label rides left), W[r] <- V.
    going right: SVD of M[(b,c,s_r),(s_l,a)], W[r] <- V*S (label rides right), W[l] <- U.
    Returns (core_l, core_r, sigma_kept)."""
    chi_l, d, chi_r = dims
    C = B.shape[1]
    full = B.T.reshape(C, chi_r, d, chi_l, d)                 # (c, b, s_r, a, s_l)
    if going_left:
        M = full.transpose(3, 0, 4, 2, 1).reshape(chi_l * C * d, d * chi_r)
        U, S, Vt = np.linalg.svd(M, full_matrices=False)
        n = truncate_spectrum(S * S, chi_max, cutoff)
        core_l = (U[:, :n] * S[:n]).reshape(chi_l, C, d, n).transpose(0, 2, 3, 1)
        core_r = Vt[:n].reshape(n, d, chi_r)
    else:
        M = full.transpose(1, 0, 2, 4, 3).reshape(chi_r * C * d, d * chi_l)
        U, S, Vt = np.linalg.svd(M, full_matrices=False)
        n = truncate_spectrum(S * S, chi_max, cutoff)
        core_r = (U[:, :n] * S[:n]).reshape(chi_r, C, d, n).transpose(3, 2, 0, 1)
        core_l = Vt[:n].reshape(n, d, chi_l).transpose(2, 1, 0)
    return np.ascontiguousarray(core_l), np.ascontiguousarray(core_r), S[:n].copy()


# --------------------------------------------------------------------------------------
# the sweep driver  (RealRealHighDimension.jl:587-890)
# --------------------------------------------------------------------------------------

def fit_sweeps(cores, phi, counts, nsweeps=1, chi_max=25, cutoff=1e-10, eta=0.01, loss="KLD",
               bbopt="TSGO", update_iters=1, rescale=(False, True), train_sep=False,
               record=None, max_bonds=None, loop=False):
    """fitMPS(W, train, test, opts) sweep loop (:726-851) + final normalize! (:852).
    cores: label on the last site, orthogonality centre there.  `record`, if a list, receives
    one dict per bond {sweep, dir, lid, loss, gradnorm, chi}.  `max_bonds` stops early
    (teacher-forcing / bounded CPU timing).  Returns the new cores."""
    cores = [c.copy() for c in cores]
    N, T, d = phi.shape
    ones = np.ones((N, 1))
    LE = construct_caches(cores, phi, going_left=True)        # :631
    RE = {}
    nb = 0

    def one_bond(j, going_left, sweep):
        nonlocal nb
        l, r = j, j + 1
        L = LE[l - 1] if l > 0 else ones
        R = RE[r + 1] if r < T - 1 else ones
        B, dims = flatten_bt(cores[l], cores[r])              # :733 / :777
        Bn, lo, gn = apply_update(B, L, R, phi[:, l], phi[:, r], counts, loss, bbopt, eta,
                                  update_iters, rescale, train_sep, loop)
        cl, cr, S = decompose_bt(Bn, dims, going_left, chi_max, cutoff)   # :756 / :798
        cores[l], cores[r] = cl, cr
        if going_left:                                        # update_caches! :759
            RE[r] = env_step_right(phi[:, r], R, cr)
        else:                                                 # :799
            LE[l] = env_step_left(phi[:, l], L, cl)
        if record is not None:
            record.append(dict(sweep=sweep, going_left=going_left, lid=l, loss=lo, gradnorm=gn,
                               chi=len(S), sigma=S))
        nb += 1
        return max_bonds is not None and nb >= max_bonds

    stop = False
    for it in range(nsweeps):
        for j in range(T - 2, -1, -1):                        # backward :731
            if one_bond(j, True, it):
                stop = True
                break
        if stop:
            break
        for j in range(T - 1):                                # forward :776
            if one_bond(j, False, it):
                stop = True
                break
        if stop:
            break
    if not stop:
        cores = normalize_mps(cores)                          # :852
    return cores


# --------------------------------------------------------------------------------------
# overlaps / classify / per-sweep metrics  (summary.jl:4-136)
# --------------------------------------------------------------------------------------

def overlaps(cores, phi):
    """contract_mps (summary.jl:4-14): yhat[i,c] = full-chain contraction, label wherever it is."""
    N, T, d = phi.shape
    E = np.ones((N, 1, 1))                                    # (i, link, class)
    for j, A in enumerate(cores):
        if A.ndim == 4:
            E = np.einsum("ia,is,askc->ikc", E[:, :, 0], phi[:, j], A, optimize=True)
        else:
            E = np.einsum("iac,is,ask->ikc", E, phi[:, j], A, optimize=True)
    return E[:, 0, :]


def classify(cores, phi, classes=None):
    """classify (summary.jl:116-136): argmax_c |yhat_c|^2, first maximum."""
    yh = overlaps(cores, phi)
    pred = np.argmax(yh * yh, axis=1)
    return pred if classes is None else np.asarray(classes)[pred]


def mse_loss_acc(cores, phi, label_idx):
    """MSE_loss_acc (summary.jl:33-70): (mse, kld, acc) averaged over samples."""
    yh = overlaps(cores, phi)
    N, C = yh.shape
    onehot = np.zeros((N, C))
    onehot[np.arange(N), label_idx] = 1.0
    mse = float(np.mean(0.5 * np.sum((yh - onehot) ** 2, axis=1)))
    kld = float(np.mean(-np.log(yh[np.arange(N), label_idx] ** 2)))
    acc = float(np.mean(np.argmax(np.abs(yh), axis=1) == label_idx))
    return mse, kld, acc


# --------------------------------------------------------------------------------------
# imputation  (Imputation/MPS_methods.jl:1-347, sampling_utils.jl:19-316, imputation.jl:48-123)
# --------------------------------------------------------------------------------------

def expand_label_index(cores):
    """utils.jl:356-370: one label-free, unit-norm MPS per class."""
    out = []
    pos = [j for j, A in enumerate(cores) if A.ndim == 4][0]
    C = cores[pos].shape[3]
    for c in range(C):
        cs = [A.copy() for A in cores]
        cs[pos] = cs[pos][..., c]
        out.append(normalize_mps(cs))
    return out


def make_grid(enc_range=(-1.0, 1.0), dx=1e-4):
    """xvals = collect(range(a, b; step=dx)) (imputation.jl:90): G = floor((b-a)/dx)+1 points,
    each the correctly rounded a + i*dx (Julia ranges use twice-precision arithmetic)."""
    a, b = enc_range
    G = int(np.floor((b - a) / dx + 1e-9)) + 1
    inv = round(1.0 / dx)
    if abs(inv * dx - 1.0) < 1e-12 and float(a * inv).is_integer():
        return (a * inv + np.arange(G)) / inv
    return a + np.arange(G) * dx


def precondition(class_cores, phi_ts, missing):
    """precondition (MPS_methods.jl:42-90) + condition_until_next! (:1-22): contract every known
    site with its encoded value; runs of known sites are absorbed into the NEXT missing core from
    the left, everything after the last missing site into the last missing core from the right
    (left-to-right matrix products, as the reference)."""
    T = len(class_cores)
    missing = sorted(int(m) for m in missing)
    known = [j for j in range(T) if j not in set(missing)]
    K = len(missing)
    cond = []
    i = 0
    idx = 0
    kn = set(known)
    while i < T and idx < K:
        if idx == K - 1:
            it = None
            while i in kn:
                Mj = np.einsum("s,asb->ab", phi_ts[i], class_cores[i])
                it = Mj if it is None else it @ Mj
                i += 1
            last = class_cores[i]
            i += 1
            it2 = None
            while i < T:                                      # all remaining sites are known
                Mj = np.einsum("s,asb->ab", phi_ts[i], class_cores[i])
                it2 = Mj if it2 is None else it2 @ Mj
                i += 1
            A = last
            if it is not None:
                A = np.einsum("xa,asb->xsb", it, A)
            if it2 is not None:
                A = np.einsum("asb,by->asy", A, it2)
            cond.append(A)
            idx += 1
        elif i in kn:
            it = None
            while i in kn:
                Mj = np.einsum("s,asb->ab", phi_ts[i], class_cores[i])
                it = Mj if it is None else it @ Mj
                i += 1
            cond.append(np.einsum("xa,asb->xsb", it, class_cores[i]))
            idx += 1
            i += 1
        else:
            cond.append(class_cores[i].copy())
            idx += 1
            i += 1
    return cond


def orthogonalize_to_first(cond):
    """ITensors orthogonalize!(mps, 1) (MPS_methods.jl:110; un-vendored: QR sweep right -> left)."""
    cond = [A.copy() for A in cond]
    for k in range(len(cond) - 1, 0, -1):
        A = cond[k]
        a, s, b = A.shape
        Qm, Rm = np.linalg.qr(A.reshape(a, s * b).T)          # (s*b, a) = Q (s*b,r) R (r,a)
        r = Qm.shape[1]
        cond[k] = Qm.T.reshape(r, s, b)
        cond[k - 1] = np.einsum("xsa,ra->xsr", cond[k - 1], Rm)
    return cond


def cond_probs(rdm, grid_enc):
    """get_conditional_probability(state::SVector, A::Matrix) (sampling_utils.jl:37-44):
    p = ||state' * rdm||^2  (the pdf uses rho^2)."""
    v = grid_enc.conj() @ rdm
    return np.real(np.sum(v * v.conj(), axis=1))


def cumtrapz_even_fast(xs, y):
    """NumericalIntegration 0.2.0 cumul_integrate(x, y, TrapezoidalEvenFast()) (un-vendored;
    published algorithm): c[0]=0; c[i]=c[i-1]+(y[i-1]+y[i]); scaled by (x[1]-x[0])/2."""
    c = np.zeros_like(y)
    c[1:] = np.cumsum(y[:-1] + y[1:])
    return (xs[1] - xs[0]) * 0.5 * c


def weighted_median(v, w):
    """StatsBase 0.34.4 `median(v, pweights(w))` = `quantile(v, w, 0.5)` (un-vendored; published algorithm restated):
    drop zero weights, sort the (value, weight) pairs, h = p*(sum(w) - w_1) + w_1 with w_1 the weight of the smallest
    value, walk the cumulative weight S_k until S_k > h and interpolate linearly between the last two values."""
    v = np.asarray(v, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    nz = w != 0
    order = np.lexsort((w[nz], v[nz]))                      # tuples sort by value, then weight
    vs, ws = v[nz][order], w[nz][order]
    wsum = float(np.sum(w))
    h = 0.5 * (wsum - ws[0]) + ws[0]
    S = np.cumsum(ws)
    k = int(np.searchsorted(S, h, side="right"))            # first k with S_k > h
    if k >= len(vs):
        return float(vs[-1])
    Sk_old = S[k - 1] if k > 0 else 0.0
    vk_old = vs[k - 1] if k > 0 else 0.0
    return float(vk_old + (h - Sk_old) / (S[k] - Sk_old) * (vs[k] - vk_old))


def orthogonalize_to_last(cond):
    """ITensors orthogonalize!(mps, length(mps)) (MPS_methods.jl:115, impute_order=:backwards): QR sweep left -> right."""
    cond = [A.copy() for A in cond]
    for k in range(len(cond) - 1):
        A = cond[k]
        a, s, b = A.shape
        Qm, Rm = np.linalg.qr(A.reshape(a * s, b))
        r = Qm.shape[1]
        cond[k] = Qm.reshape(a, s, r)
        cond[k + 1] = np.einsum("rb,bsy->rsy", Rm, cond[k + 1])
    return cond


def impute_series_ex(class_cores, x_scaled, missing, grid, grid_enc, d, basis="legendre_no_norm",
                     method="median", uniforms=None, max_jump=None, impute_order="forwards", get_err=False,
                     rejection_threshold=None, max_trials=10):
    """impute_median / impute_mean / impute_mode / impute_ITS (MPS_methods.jl:201-347) -> impute_at! (:93-180) on one
    series already in the encoding range with the missing entries filled (imputation.jl:290-291).
    `impute_order`: "forwards" (orthogonalize!(mps, 1), walk first -> last missing site) or "backwards" (:102-118).
    `get_err`: the method's error bar per imputed site -- WMAD for the median (get_wmad, sampling_utils.jl:192-196),
    the standard deviation for the mean (get_std, :89-97), the WMAD for ITS with rejection (:295-311), 0 otherwise.
    `rejection_threshold` (ITS): None = :none; else up to `max_trials` draws per site, accepted when
    |x - median| < threshold * WMAD (the last draw stands if none is accepted); `uniforms` is then the flat stream of
    rand(rng) values consumed in order (the reference shares one MersenneTwister over sites and trajectories).
    Returns (x_out (T,), grid_index (K,), errs (T,), draws_used)."""
    T = len(class_cores)
    missing = sorted(int(m) for m in missing)
    x_out = np.array(x_scaled, dtype=np.float64).copy()
    errs = np.zeros(T)
    # data-driven / time-dependent encodings (imputation.jl:92-100: xvals_enc[site]): `basis` is then a callable
    # basis(site, x) -> (..., d) and `grid_enc` holds one (G, d) block per site
    site_enc = callable(basis)
    grid_enc_all = np.asarray(grid_enc)
    enc1 = (lambda j, xv: np.asarray(basis(j, np.asarray(xv, dtype=np.float64))).reshape(-1)) if site_enc else None
    phi_ts = np.stack([enc1(j, x_out[j]) for j in range(T)]) if site_enc else encode(x_out, d, basis)
    cond = precondition(class_cores, phi_ts, missing)
    K = len(missing)
    fwd = impute_order == "forwards"
    if fwd:
        cond = orthogonalize_to_first(cond)
        A = cond[0][0]                                        # (d, chi)
        order = list(range(K))
        x_prev = x_out[missing[0] - 1] if missing[0] > 0 else None            # MPS_methods.jl:136-138
    elif impute_order == "backwards":
        cond = orthogonalize_to_last(cond)
        A = cond[-1][:, :, 0].T                               # (d, chi): site first, link to the rest second
        order = list(range(K - 1, -1, -1))
        x_prev = x_out[missing[-1] + 1] if missing[-1] + 1 < T else None      # :139-141
    else:
        raise ValueError('impute_order must be either "forwards" or "backwards"')
    idxs = np.zeros(K, dtype=np.int64)
    dxm = float(np.mean(np.abs(np.diff(grid))))
    ucur = 0
    for ii, k in enumerate(order):
        rdm = A @ A.conj().T                                  # :152
        grid_enc = grid_enc_all[missing[k]] if grid_enc_all.ndim == 3 else grid_enc_all
        p = cond_probs(rdm, grid_enc)
        err = 0.0
        if method == "median":                                # sampling_utils.jl:162-199
            cdf = cumtrapz_even_fast(grid, p)
            Z = cdf[-1]
            g = int(np.argmin(np.abs(cdf / Z - 0.5)))
            xv, st = grid[g], grid_enc[g] / np.sqrt(Z)
            if get_err:
                err = weighted_median(np.abs(grid - xv), p / Z)
        elif method == "ITS" and rejection_threshold is None:   # :263-290 (no rejection)
            cdf = cumtrapz_even_fast(grid, p)
            cdf = cdf / cdf[-1]
            Z = cdf[-1]
            u = uniforms[ucur]
            ucur += 1
            g = int(np.argmin(np.abs(cdf / Z - u)))
            xv, st = grid[g], grid_enc[g] / np.sqrt(Z)
        elif method == "ITS":                                 # :291-311 (rejection by WMAD)
            cdf = cumtrapz_even_fast(grid, p)
            Zr = cdf[-1]
            cdf = cdf / Zr
            gm = int(np.argmin(np.abs(cdf - 0.5)))
            wmad = weighted_median(np.abs(grid - grid[gm]), p / Zr)
            Z = cdf[-1]
            g = gm
            for _ in range(max_trials):
                u = uniforms[ucur]
                ucur += 1
                g = int(np.argmin(np.abs(cdf / Z - u)))
                if abs(grid[g] - grid[gm]) < rejection_threshold * wmad:
                    break
            xv, st = grid[g], grid_enc[g] / np.sqrt(Z)
            err = wmad
        elif method == "mode":                                # :104-158
            if x_prev is None or max_jump is None:
                g = int(np.argmax(p))
            else:
                perm = np.argsort(-p, kind="stable")
                ok = np.abs(grid[perm] - x_prev) <= max_jump
                g = int(perm[np.argmax(ok)]) if ok.any() else int(perm[0])
            xv, st = grid[g], grid_enc[g]
        elif method == "mean":                                # :64-101
            Z = (grid[1] - grid[0]) * (0.5 * (p[0] + p[-1]) + np.sum(p[1:-1]))
            xv = float(np.sum(grid * p) * dxm / Z)
            st = (enc1(missing[k], xv) if site_enc else encode(np.array(xv), d, basis)) / np.sqrt(Z)
            g = -1
            if get_err:
                err = float(np.sqrt(np.sum((grid - xv) ** 2 * p) * dxm / Z))
        else:
            raise ValueError(method)
        idxs[k] = g
        x_out[missing[k]] = xv
        errs[missing[k]] = err
        x_prev = xv
        if ii != K - 1:                                       # MPS_methods.jl:161-168, norm=false
            v = st.conj() @ A
            if fwd:
                A = np.einsum("a,asb->sb", v, cond[k + 1])
            else:
                A = np.einsum("asb,b->sa", cond[k - 1], v)
    return x_out, idxs, errs, ucur


def impute_series(class_cores, x_scaled, missing, grid, grid_enc, d, basis="legendre_no_norm",
                  method="median", uniforms=None, max_jump=None):
    """Forward-order imputation without error bars: (x_out (T,), grid_index (K,)).  See impute_series_ex."""
    x_out, idxs, _, _ = impute_series_ex(class_cores, x_scaled, missing, grid, grid_enc, d, basis, method, uniforms,
                                         max_jump)
    return x_out, idxs


# --------------------------------------------------------------------------------------
# synthetic inputs of the BASELINE shapes (toy_data.jl:53-85,
# missing_data_mechanisms.jl:146-153) -- own seeded RNG, not Julia's stream
# --------------------------------------------------------------------------------------

def trendy_sine(T, n, period, slope_choices, sigma, rng):
    """x_t = sin(2*pi*t/tau + psi) + m*t/T + sigma*n_t, t=1..T (toy_data.jl:74-78)."""
    tau = rng.uniform(period[0], period[1], size=n)
    m = rng.choice(np.asarray(slope_choices, dtype=np.float64), size=n)
    psi = rng.uniform(0.0, 2 * np.pi, size=n)
    t = np.arange(1, T + 1, dtype=np.float64)
    X = np.sin(2 * np.pi / tau[:, None] * t[None, :] + psi[:, None]) + m[:, None] * t[None, :] / T
    X += sigma * rng.standard_normal((n, T))
    return X


def synthetic_two_class(N, T, seed, periods=((12.0, 15.0), (16.0, 19.0)), sigma=0.1):
    rng = np.random.default_rng(seed)
    n0 = N // 2
    X0 = trendy_sine(T, n0, periods[0], (-3.0, 0.0, 3.0), sigma, rng)
    X1 = trendy_sine(T, N - n0, periods[1], (-3.0, 0.0, 3.0), sigma, rng)
    X = np.concatenate([X0, X1], axis=0)
    y = np.concatenate([np.zeros(n0, dtype=np.int64), np.ones(N - n0, dtype=np.int64)])
    return X, y
