"""Host-side initialisers of the data-driven / time-dependent encodings (reference src/Encodings/bases.jl:135-397,
splitbases.jl:1-110).  They run ONCE per fit on the (already normalised) training set and stay on the CPU in the
reference as well (`opts.encoding.init`, encodings.jl:112-120); what they return here is the per-site coefficient
table that K1 evaluates on the device (csrc/encode_table.cu, mpst_set_encoding_table).

The Julia shim keeps the reference's own init functions (KernelDensity / Interpolations / Integrals) and only packs
their results into the same tables; the numpy restatements below stand in for them where Julia is absent.  They follow
the published algorithms of the un-vendored packages (KernelDensity 0.6: Silverman bandwidth, linear binning on 2048
points, FFT convolution with the Gaussian; Interpolations: quadratic B-spline, `Line(OnGrid())` boundary) -- parity of
the *tables* with Julia's is unpinned, parity of the *encoding given a table* is what the tests check."""
import numpy as np

TABLE_LEGENDRE_PROJ, TABLE_SAHAND_LEGENDRE, TABLE_SPLIT = 101, 102, 103
AUX_IDS = {"legendre_no_norm": 0, "legendre": 0, "legendre_norm": 1, "uniform": 5}


# ---- KernelDensity.jl kde(data) -------------------------------------------------------------------------------------
def kde(data, bandwidth=None, npoints=2048):
    """(x grid, density): kde(data; boundary=(lo-4h, hi+4h), npoints=2048, kernel=Normal, bandwidth=Silverman)."""
    data = np.asarray(data, dtype=np.float64)
    n = data.size
    if bandwidth is None:
        if n <= 1:
            bandwidth = 0.9
        else:
            q25, q75 = np.quantile(data, [0.25, 0.75])
            width = min(np.std(data, ddof=1), (q75 - q25) / 1.34)
            if width == 0.0:
                width = 1.0 if np.all(data == 0) else abs(np.mean(data))
            bandwidth = 0.9 * width * n ** (-0.2)
    lo, hi = data.min() - 4.0 * bandwidth, data.max() + 4.0 * bandwidth
    x = lo + (hi - lo) / (npoints - 1) * np.arange(npoints)
    s = x[1] - x[0]
    grid = np.zeros(npoints)
    k = np.searchsorted(x, data, side="left")                 # searchsortedfirst (0-based: midpoints[k] >= x)
    ok = (k >= 1) & (k <= npoints - 1)
    k, xd = k[ok], data[ok]
    ainc = 1.0 / (n * s * s)
    np.add.at(grid, k - 1, (x[k] - xd) * ainc)
    np.add.at(grid, k, (xd - x[k - 1]) * ainc)
    ft = np.fft.rfft(grid)
    c = -2.0 * np.pi / (s * npoints)
    t = np.arange(ft.size) * c
    ft = ft * np.exp(-0.5 * (bandwidth * t) ** 2)             # characteristic function of Normal(0, bandwidth)
    dens = np.maximum(np.fft.irfft(ft, npoints), 0.0)
    return x, dens


def bspline_quadratic_coeffs(y):
    """Interpolations.jl prefilter of BSpline(Quadratic(Line(OnGrid()))): padded coefficients c[0..n+1] with
    c[i-1]/8 + 3 c[i]/4 + c[i+1]/8 = y[i] (i = 1..n) and zero second derivative at the two edge knots."""
    from scipy.linalg import solve_banded
    y = np.asarray(y, dtype=np.float64)
    n = y.size
    m = n + 2
    ab = np.zeros((5, m))                                     # bands u2, u1, diag, l1, l2
    ab[2, 1:-1] = 0.75
    ab[1, 2:] = 0.125                                         # A[i, i+1] stored at ab[1, i+1]
    ab[3, :-2] = 0.125                                        # A[i, i-1] stored at ab[3, i-1]
    ab[2, 0], ab[1, 1], ab[0, 2] = 1.0, -2.0, 1.0             # c0 - 2 c1 + c2 = 0
    ab[2, -1], ab[3, -2], ab[4, -3] = 1.0, -2.0, 1.0          # c_{n-1} - 2 c_n + c_{n+1} = 0
    rhs = np.concatenate([[0.0], y, [0.0]])
    return solve_banded((2, 2), ab, rhs)


def interp_kde_pdf(x, grid_x, coeffs):
    """pdf(InterpKDE(kde), x): the quadratic B-spline at x, zero outside the grid."""
    x = np.asarray(x, dtype=np.float64)
    n = grid_x.size
    h = grid_x[1] - grid_x[0]
    t = (x - grid_x[0]) / h + 1.0
    inside = (t >= 1.0) & (t <= n)
    tr = np.rint(np.where(inside, t, 1.0))
    i = tr.astype(np.int64)
    dx = np.where(inside, t, 1.0) - tr
    v = coeffs[i - 1] * 0.5 * (dx - 0.5) ** 2 + coeffs[i] * (0.75 - dx * dx) + coeffs[i + 1] * 0.5 * (dx + 0.5) ** 2
    return np.where(inside, v, 0.0)


def _trapz(y, x):
    return float(np.sum(0.5 * (y[1:] + y[:-1]) * np.diff(x)))


def _legendre_normalised(x, L):
    """P_l(x) sqrt((2l+1)/2), l = 0..L-1, rows = orders."""
    x = np.asarray(x, dtype=np.float64)
    P = np.empty((L, x.size))
    P[0] = 1.0
    if L > 1:
        P[1] = x
    for l in range(2, L):
        P[l] = ((2 * l - 1) * x * P[l - 1] - (l - 1) * P[l - 2]) / l
    return P * np.sqrt((2 * np.arange(L) + 1) / 2.0)[:, None]


# ---- projected Legendre (bases.jl:381-397, 346-356) -------------------------------------------------------------------
def project_legendre(X_norm_TxN, d, norm=False, enc_range=(-1.0, 1.0), max_series_terms=None):
    """Per time point: the d series terms with the largest |<wf, P_l>| of the KDE wavefunction of that time point's
    values.  Returns the device table (kind, n_sites, ip, dp) and the chosen orders (T, d)."""
    X = np.asarray(X_norm_TxN, dtype=np.float64)
    T = X.shape[0]
    L = max_series_terms or 7 * d
    a, b = enc_range
    orders = np.zeros((T, d), dtype=np.int64)
    for t in range(T):
        xs = X[t][(X[t] >= a) & (X[t] <= b)]
        ns = max(200, 2 * xs.size)
        xg = np.linspace(-1.0, 1.0, ns)
        gx, dens = kde(xs)
        wf = np.sqrt(np.maximum(interp_kde_pdf(xg, gx, bspline_quadratic_coeffs(dens)), 0.0))
        basis = _legendre_normalised(xg, L)
        coeffs = np.array([_trapz(wf * basis[l], xg) for l in range(L)])
        # partialsortperm(abs2.(coeffs), 1:d; rev=true) returns 1-based positions in the list P_0, P_1, ...; the encoder
        # then uses each position AS the polynomial order (legendre(x, d, nds) = Pl(x, d), bases.jl:97): position i
        # was the coefficient of P_{i-1} but P_i is evaluated.  Restated literally (so P_0 is never part of the basis).
        orders[t] = np.argsort(-np.abs(coeffs) ** 2, kind="stable")[:d] + 1
    return legendre_proj_table(orders, norm), orders


def legendre_proj_table(orders, norm=False):
    orders = np.asarray(orders, dtype=np.int64)
    T, d = orders.shape
    Lmax = int(orders.max())
    ni = d + Lmax + 1
    ip = -np.ones((T, ni), dtype=np.int32)
    dp = np.zeros((T, 2))
    for t in range(T):
        ip[t, :d] = orders[t]
        L = int(orders[t].max())
        for k, l in enumerate(orders[t]):
            ip[t, d + l] = k
        dmax = int(orders[t].max())
        # bases.jl:101-105: ls /= sqrt(Pl(1, dmax; normalized) * dmax)
        dp[t, 0] = 1.0 / np.sqrt(np.sqrt((2 * dmax + 1) / 2.0) * dmax) if (norm and dmax > 0) else 1.0
        dp[t, 1] = L
    return TABLE_LEGENDRE_PROJ, T, ip, dp


# ---- Sahand-Legendre (bases.jl:168-215, 257-342) --------------------------------------------------------------------------
def sahand_legendre_coeffs(xs, f0, d):
    """bases.jl:168-215: polynomials c_n(x) orthonormal under the weight f0^2, built row by row from the moment matrix."""
    N = d - 1
    cV = np.zeros((N + 1, N + 1))
    cV[0, 0] = 1.0
    M = np.array([[_trapz(xs ** (i + j) * f0 ** 2, xs) for j in range(N + 1)] for i in range(N + 1)])
    for n in range(1, N + 1):
        if n == 1:
            cV[1, 0] = 1.0
            cV[1, 1] = -1.0 / M[1, 0]
            nrm = cV[1, :2] @ M[:2, :2] @ cV[1, :2]
            cV[1] /= np.sqrt(nrm)
        else:
            cvt = cV[:n, :n] @ M[0, :n]
            A = cV[:n, :n] @ M[1:n + 1, :n].T
            sol = np.linalg.solve(A, -cvt)
            cV[n, 0] = 1.0
            cV[n, 1:n + 1] = sol
            nrm = cV[n, :n + 1] @ M[:n + 1, :n + 1] @ cV[n, :n + 1]
            cV[n] /= np.sqrt(nrm)
    return cV


def remove_zeros(xs, f0):
    """bases.jl:257-283: floor the wavefunction at the smallest value above 1 % of its maximum, then divide it by the
    integral of its square.  Returns (f0 modified, minval, norm)."""
    f0 = f0.copy()
    tol = np.max(np.abs(f0)) * 1e-2
    bad = np.abs(f0) <= tol
    if bad.all():
        return f0, 0.0, 1.0
    minval = float(np.min(np.abs(f0[~bad])))
    f0[bad] = minval
    nrm = _trapz(f0 ** 2, xs)
    return f0 / nrm, minval, nrm


def init_sahand_legendre(X_norm_TxN, d, time_dependent=True, enc_range=(-1.0, 1.0), max_samples=None):
    """init_sahand_legendre(_time_dependent) (bases.jl:294-342) -> device table."""
    X = np.asarray(X_norm_TxN, dtype=np.float64)
    T = X.shape[0]
    a, b = enc_range
    rows = [X[t] for t in range(T)] if time_dependent else [X.reshape(-1)]
    ns = max_samples or max(200, T)
    npts = 2048
    ip = np.zeros((len(rows), 2), dtype=np.int32)
    dp = np.zeros((len(rows), 4 + d * d + npts + 2))
    xg = np.linspace(a, b, ns)
    for t, row in enumerate(rows):
        xs = row[(row >= a) & (row <= b)]
        ip[t, 0] = npts
        dp[t, 1] = 1.0
        dp[t, 3] = 1.0
        if xs.size == 0:
            continue
        gx, dens = kde(xs, npoints=npts)
        cs = bspline_quadratic_coeffs(dens)
        f0 = np.sqrt(np.maximum(interp_kde_pdf(xg, gx, cs), 0.0))
        f0n, minx, scale = remove_zeros(xg, f0)
        if minx == 0.0:
            continue
        ip[t, 1] = 1
        dp[t, 0], dp[t, 1], dp[t, 2], dp[t, 3] = gx[0], gx[1] - gx[0], minx, scale
        dp[t, 4:4 + d * d] = sahand_legendre_coeffs(xg, f0n, d).reshape(-1)
        dp[t, 4 + d * d:] = cs
    return TABLE_SAHAND_LEGENDRE, len(rows), ip, dp


# ---- split bases (splitbases.jl:1-110) -----------------------------------------------------------------------------------
def unif_split(nbins, a, b):
    dx = (b - a) / nbins
    return a + dx * np.arange(nbins + 1)                      # collect(a:dx:b)


def hist_split(samples, nbins, a, b):
    """splitbases.jl:58-90, one time point."""
    samples = np.asarray(samples, dtype=np.float64)
    npts = samples.size
    bin_pts = int(round(npts / nbins))
    if bin_pts == 0:
        bin_pts = 1
    bins = np.full(nbins + 1, float(a))
    j = 1
    ds = np.sort(samples[(samples >= a) & (samples <= b)])
    for i, x in enumerate(ds, start=1):
        if i % bin_pts == 0 and i < npts:
            if j == nbins:
                break
            bins[j] = (x + ds[i]) / 2 if i < ds.size else x
            j += 1
    if j <= nbins - 1:
        bins[bins == a] = b
        bins[0] = a
    bins[-1] = b
    return bins


def split_table(X_norm_TxN, d, aux_basis_dim, aux="uniform", method="hist", enc_range=(0.0, 1.0)):
    """split_init (splitbases.jl:13-52) for a data-independent auxiliary basis -> device table."""
    if d % aux_basis_dim:
        raise ValueError(f"The auxilliary basis dimension ({aux_basis_dim}) must evenly divide the total feature dimension ({d})")
    X = np.asarray(X_norm_TxN, dtype=np.float64)
    nbins = d // aux_basis_dim
    a, b = enc_range
    if method == "unif":
        rows = [unif_split(nbins, a, b)]
    else:
        rows = [hist_split(X[t], nbins, a, b) for t in range(X.shape[0])]
    ip = np.tile(np.array([nbins, aux_basis_dim, AUX_IDS[aux.lower()]], dtype=np.int32), (len(rows), 1))
    return TABLE_SPLIT, len(rows), ip, np.asarray(rows, dtype=np.float64)
