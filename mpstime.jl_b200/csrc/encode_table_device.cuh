// Device-side evaluation of the table encodings (see encode_table.cu for the table layouts): shared by K1
// (encode_table.cu) and K8 (impute.cu, which encodes known / imputed values of one site on the fly).
#pragma once
#include "mpst_common.cuh"
#include "encode_device.cuh"

namespace {

// ---- projected Legendre: d selected orders out of the first L+1 normalised polynomials --------------------------------
// ip[0 .. d)        : order (0-based) of output k          (unused by the kernel, kept for read-back / debugging)
// ip[d .. d+L+1)    : inverse map order -> output slot or -1
// dp[0]             : scale (1 for "no norm", 1/sqrt(Pl(1, dmax) * dmax) for the normalised variant, bases.jl:101-105)
// dp[1]             : L = highest order needed
__device__ __forceinline__ void legendre_proj_point(double x, int d, const int* __restrict__ ip, const double* __restrict__ dp,
                                                    double* v) {
    const double scale = dp[0];
    const int L = (int)dp[1];
    const int* inv = ip + d;
    double pm = 1.0, p = x;
    for (int k = 0; k < d; k++) v[k] = 0.0;
    for (int l = 0; l <= L; l++) {
        double pl;
        if (l == 0) pl = 1.0;
        else if (l == 1) pl = x;
        else { pl = ((double)(2 * l - 1) * x * p - (double)(l - 1) * pm) / (double)l; pm = p; p = pl; }
        const int slot = inv[l];
        if (slot >= 0) v[slot] = sqrt((double)(2 * l + 1) * 0.5) * pl * scale;
    }
}

// ---- Sahand-Legendre: orthonormal polynomials under the data density f0^2, times f0 -----------------------------------
// dp[0] = x_1 (first KDE grid point), dp[1] = h (KDE grid spacing), dp[2] = minx, dp[3] = scale,
// dp[4 .. 4 + d*d) = cVecs[n][i] (coefficient of x^i in basis function n),
// dp[4 + d*d ...)  = quadratic B-spline coefficients c[0 .. npts+1] of the density (ghost coefficients at both ends)
// ip[0] = npts (KDE grid points); ip[1] = 1 if the site has data (else every output is 0, bases.jl:318-327)
// pdf(kde, x): KernelDensity.InterpKDE = Interpolations BSpline(Quadratic(Line(OnGrid()))) scaled to the grid, zero outside
__device__ __forceinline__ void sahand_legendre_point(double x, int d, const int* __restrict__ ip, const double* __restrict__ dp,
                                                      double* v) {
    if (!ip[1]) { for (int n = 0; n < d; n++) v[n] = 0.0; return; }
    const int npts = ip[0];
    const double x1 = dp[0], h = dp[1], minx = dp[2], scale = dp[3];
    const double* cv = dp + 4;
    const double* cs = cv + d * d;
    double pdf = 0.0;
    const double t = (x - x1) / h + 1.0;                 // 1-based fractional grid index
    if (t >= 1.0 && t <= (double)npts) {
        const double tr = rint(t);                       // Julia round(): ties to even, like rint
        const int i = (int)tr;
        const double dx = t - tr;
        const double wm = 0.5 * (dx - 0.5) * (dx - 0.5), w0 = 0.75 - dx * dx, wp = 0.5 * (dx + 0.5) * (dx + 0.5);
        pdf = cs[i - 1] * wm + cs[i] * w0 + cs[i + 1] * wp;
    }
    const double f0 = fmax(sqrt(fmax(pdf, 0.0)), minx);
    for (int n = 0; n < d; n++) {
        // sum(c * x^(i-1)) in the reference's order (bases.jl:115): ascending powers
        double s = 0.0, xp = 1.0;
        for (int i = 0; i < d; i++) { s += cv[n * d + i] * xp; xp *= x; }
        v[n] = s * f0 / scale;
    }
}

// ---- split basis: an auxiliary basis of dimension aux_dim copied into every bin of [a, b] ----------------------------------
// ip[0] = nbins, ip[1] = aux_dim, ip[2] = auxiliary basis id (data-independent real bases only); dp[0 .. nbins] = bin edges
__device__ __forceinline__ void split_point(double x, int d, const int* __restrict__ ip, const double* __restrict__ dp, double* v) {
    const int nbins = ip[0], ad = ip[1], aux = ip[2];
    const double a = dp[0], b = dp[nbins];
    const double scale = b - a;
    for (int i = 0; i < nbins; i++) {
        const double dxb = dp[i + 1] - dp[i];
        const double lb = i == 0 ? 1.0 : 0.5, rb = i == nbins - 1 ? 1.0 : 0.5;
        const double xprop = scale * (x - dp[i]) / dxb;              // splitbases.jl:128
        const double r = xprop / scale - 0.5;
        const double sel = r == -0.5 ? lb : (r == 0.5 ? rb : ((r >= -0.5 && r <= 0.5) ? 1.0 : 0.0));      // rect(), :96-109
        double* o = v + i * ad;
        if (sel == 0.0) { for (int k = 0; k < ad; k++) o[k] = 0.0; continue; }
        double w[MPST_MAX_D];
        const double xa = a + xprop;
        if (aux == MPST_BASIS_LEGENDRE_NO_NORM) legendre_point(xa, ad, false, w);
        else if (aux == MPST_BASIS_LEGENDRE_NORM) legendre_point(xa, ad, true, w);
        else for (int k = 0; k < ad; k++) w[k] = 1.0 / (double)ad;
        for (int k = 0; k < ad; k++) o[k] = sel * w[k];
    }
}

// one point of table encoding `kind` (MPST_BASIS_TABLE_*) with the tables of one site
__device__ __forceinline__ void table_point(int kind, double x, int d, const int* __restrict__ ip, const double* __restrict__ dp, double* v) {
    if (kind == MPST_BASIS_TABLE_LEGENDRE_PROJ) legendre_proj_point(x, d, ip, dp, v);
    else if (kind == MPST_BASIS_TABLE_SAHAND_LEGENDRE) sahand_legendre_point(x, d, ip, dp, v);
    else split_point(x, d, ip, dp, v);
}
}  // namespace
