// K8: batched MPS imputation (replaces, per instance, get_predictions -> precondition -> impute_at!
// -> get_{median,mean,mode,sample}_from_rdm; reference Imputation/imputation.jl:264-410,
// Imputation/MPS_methods.jl:42-180, Imputation/sampling_utils.jl:19-316).
//
// One instance per CTA (persistent grid over the batch, no inter-instance communication).
// The reference conditions the class MPS on the known sites, brings the conditioned cores to
// right-canonical form (orthogonalize!, a QR sweep) and then walks the missing sites left to right.
// Here the same reduced density matrices are obtained without QR: a backward pass carries the right
// Gram matrix  G_j = sum over everything right of site j  (chi x chi, rank 1 until the last missing site),
//     known site  : G <- M G M^T,   M = sum_s phi(x_j)[s] A_j[:, s, :]
//     missing site: G <- sum_s A_j[:, s, :] G A_j[:, s, :]^T
// and stores it for every missing site; the forward pass carries the left vector v, forms
// A = v . A_j (d x chi), rho = A G A^T (d x d; equal to the reference's A A' after orthogonalisation),
// evaluates the conditional pdf p[g] = ||rho Phi_g||^2 on the grid (sampling_utils.jl:37-44),
// the cumulative trapezoid (NumericalIntegration cumul_integrate, TrapezoidalEvenFast) and picks
// argmin |cdf/Z - 1/2| (median), argmin |cdf/Z - u| (ITS), argmax p (mode) or E[x] (mean); the chosen state
// is projected in (v <- state . A) and the walk continues.  Results are scale invariant, so v
// and G are renormalised freely to stay in range.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include "mpst_common.cuh"
#include "encode_device.cuh"

namespace {
constexpr int NT = 256;

struct ImpParams {
    const double* cores;        // per site [s][a][b]
    const int64_t* core_off;    // [T]
    const int* chi;             // [T+1] link dims (chi[0] = chi[T] = 1)
    const double* X;            // [n][T] scaled series, missing filled
    const uint8_t* mask;        // [n][T]
    const double* grid;         // [G]
    const double* genc;         // [G][d]
    const double* uniforms;     // [n][ntraj][Kmax] or null
    double* out;                // [n][ntraj][T]
    double* gr_scratch;         // [grid][Kmax][chimax^2]
    double* p_scratch;          // [grid][G]
    int T, d, G, ntraj, Kmax, chimax, method, basis;
    int64_t n;
    double max_jump;
};

__device__ __forceinline__ void encode_any(int basis, double x, int d, double* v) {
    if (basis == MPST_BASIS_LEGENDRE_NO_NORM) encode_point<MPST_BASIS_LEGENDRE_NO_NORM>(x, d, v);
    else if (basis == MPST_BASIS_LEGENDRE_NORM) encode_point<MPST_BASIS_LEGENDRE_NORM>(x, d, v);
    else encode_point<MPST_BASIS_UNIFORM>(x, d, v);
}

// C[n x n2] = A[n x k] * B[k x n2] (or B^T when TB), all in shared memory with pitch ld.
template <bool TB, bool ACC>
__device__ __forceinline__ void smem_matmul(double* __restrict__ Cm, const double* __restrict__ A,
                                            const double* __restrict__ B, int n, int k, int n2, int ld) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int i0 = 0; i0 < n; i0 += 64)
        for (int j0 = 0; j0 < n2; j0 += 64) {
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
            for (int kk = 0; kk < k; kk++) {
                double av[4], bv[4];
#pragma unroll
                for (int a = 0; a < 4; a++) { const int i = i0 + ty + 16 * a; av[a] = i < n ? A[i * ld + kk] : 0.0; }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int j = j0 + tx + 16 * b;
                    bv[b] = j < n2 ? (TB ? B[j * ld + kk] : B[kk * ld + j]) : 0.0;
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[a][b] += av[a] * bv[b];
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
                    if (i < n && j < n2) { if (ACC) Cm[i * ld + j] += acc[a][b]; else Cm[i * ld + j] = acc[a][b]; }
                }
        }
}

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < NT / 32; w++) t += sh[w];
    __syncthreads();
    return t;
}
__device__ __forceinline__ double block_reduce_max(double v, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = sh[0];
    for (int w = 1; w < NT / 32; w++) t = fmax(t, sh[w]);
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(NT, 1) impute_kernel(ImpParams P) {
    extern __shared__ double sm[];
    const int ld = P.chimax + 1;                       // odd-ish pitch against bank conflicts
    const int msz = P.chimax * ld;
    double* Gm = sm;                                   // current right Gram
    double* Gn = Gm + msz;                             // next
    double* As = Gn + msz;                             // staged A_s / M
    double* T1 = As + msz;                             // temp product
    double* vec = T1 + msz;                            // [chimax] left vector / right vector
    double* vec2 = vec + P.chimax;
    double* Ad = vec2 + P.chimax;                      // [d][ld]
    double* Td = Ad + P.d * ld;                        // [d][ld]
    double* rho = Td + P.d * ld;                       // [d][d]
    double* rho2 = rho + P.d * P.d;                    // [d][d]
    double* phi = rho2 + P.d * P.d;                    // [d]
    double* red = phi + MPST_MAX_D;                    // [32] reductions
    double* scr = red + 32;                            // [3*NT + 16] scan / arg-reduction scratch
    __shared__ int s_misc[4];
    const int tid = threadIdx.x;
    const int T = P.T, d = P.d, G = P.G;
    double* GR = P.gr_scratch + (size_t)blockIdx.x * P.Kmax * P.chimax * P.chimax;
    double* pbuf = P.p_scratch + (size_t)blockIdx.x * G;

    for (int64_t inst = blockIdx.x; inst < P.n; inst += gridDim.x) {
        const double* x = P.X + inst * T;
        const uint8_t* mk = P.mask + inst * T;
        // first / last missing site and count
        if (tid == 0) {
            int first = -1, last = -1, K = 0;
            for (int j = 0; j < T; j++) if (mk[j]) { if (first < 0) first = j; last = j; K++; }
            s_misc[0] = first; s_misc[1] = last; s_misc[2] = K;
        }
        __syncthreads();
        const int first = s_misc[0], last = s_misc[1], K = s_misc[2];
        for (int tr = 0; tr < P.ntraj; tr++)
            for (int j = tid; j < T; j += NT) P.out[(inst * P.ntraj + tr) * T + j] = x[j];
        if (K == 0) { __syncthreads(); continue; }

        // ---------------- backward pass: right Gram matrices ------------------------------------
        // right of the last missing site everything is known: rank-1 Gram r r^T, carry the vector r
        for (int b = tid; b < P.chimax; b += NT) vec[b] = (b == 0) ? 1.0 : 0.0;
        __syncthreads();
        for (int j = T - 1; j > last; j--) {
            const int cl = P.chi[j], cr = P.chi[j + 1];
            if (tid == 0) encode_any(P.basis, x[j], d, phi);
            __syncthreads();
            const double* A = P.cores + P.core_off[j];           // [s][a][b]
            for (int a = tid; a < cl; a += NT) {
                double acc = 0.0;
                for (int s = 0; s < d; s++) {
                    const double* row = A + ((size_t)s * cl + a) * cr;
                    double t = 0.0;
                    for (int b = 0; b < cr; b++) t += row[b] * vec[b];
                    acc += phi[s] * t;
                }
                vec2[a] = acc;
            }
            __syncthreads();
            double nn = 0.0;
            for (int a = tid; a < cl; a += NT) nn += vec2[a] * vec2[a];
            nn = block_reduce_sum(nn, red);
            const double sc = nn > 0.0 ? rsqrt(nn) : 1.0;
            for (int a = tid; a < P.chimax; a += NT) vec[a] = a < cl ? vec2[a] * sc : 0.0;
            __syncthreads();
        }
        {   // G = r r^T  (dimension chi[last+1])
            const int cr = P.chi[last + 1];
            for (int e = tid; e < cr * cr; e += NT) Gm[(e / cr) * ld + (e % cr)] = vec[e / cr] * vec[e % cr];
            __syncthreads();
        }
        int kidx = K - 1;
        for (int j = last; j >= first; j--) {
            const int cl = P.chi[j], cr = P.chi[j + 1];
            const double* A = P.cores + P.core_off[j];
            if (mk[j]) {
                // store the Gram this missing site sees
                double* dst = GR + (size_t)kidx * P.chimax * P.chimax;
                for (int e = tid; e < cr * cr; e += NT) dst[e] = Gm[(e / cr) * ld + (e % cr)];
                kidx--;
                if (j == first) break;
                for (int e = tid; e < cl * cl; e += NT) Gn[(e / cl) * ld + (e % cl)] = 0.0;
                for (int s = 0; s < d; s++) {
                    __syncthreads();
                    for (int e = tid; e < cl * cr; e += NT) As[(e / cr) * ld + (e % cr)] = A[(size_t)s * cl * cr + e];
                    __syncthreads();
                    smem_matmul<false, false>(T1, As, Gm, cl, cr, cr, ld);      // T1 = A_s G
                    __syncthreads();
                    smem_matmul<true, true>(Gn, T1, As, cl, cr, cl, ld);        // Gn += T1 A_s^T
                }
            } else {
                if (tid == 0) encode_any(P.basis, x[j], d, phi);
                __syncthreads();
                for (int e = tid; e < cl * cr; e += NT) {
                    double acc = 0.0;
                    for (int s = 0; s < d; s++) acc += phi[s] * A[(size_t)s * cl * cr + e];
                    As[(e / cr) * ld + (e % cr)] = acc;                         // M
                }
                __syncthreads();
                smem_matmul<false, false>(T1, As, Gm, cl, cr, cr, ld);          // T1 = M G
                __syncthreads();
                smem_matmul<true, false>(Gn, T1, As, cl, cr, cl, ld);           // Gn = T1 M^T
            }
            __syncthreads();
            // renormalise (scale invariance) and swap
            double mx = 0.0;
            for (int e = tid; e < cl * cl; e += NT) mx = fmax(mx, fabs(Gn[(e / cl) * ld + (e % cl)]));
            mx = block_reduce_max(mx, red);
            const double sc = mx > 0.0 ? 1.0 / mx : 1.0;
            for (int e = tid; e < cl * cl; e += NT) Gm[(e / cl) * ld + (e % cl)] = Gn[(e / cl) * ld + (e % cl)] * sc;
            __syncthreads();
        }
        __syncthreads();

        // ---------------- forward pass(es) ---------------------------------------------------------
        for (int tr = 0; tr < P.ntraj; tr++) {
            double* xo = P.out + (inst * P.ntraj + tr) * T;
            for (int a = tid; a < P.chimax; a += NT) vec[a] = (a == 0) ? 1.0 : 0.0;
            __syncthreads();
            int k = 0;
            double x_prev = 0.0;
            bool have_prev = false;
            for (int j = 0; j <= last; j++) {
                const int cl = P.chi[j], cr = P.chi[j + 1];
                const double* A = P.cores + P.core_off[j];
                if (!mk[j]) {
                    if (tid == 0) encode_any(P.basis, x[j], d, phi);
                    __syncthreads();
                    for (int b = tid; b < cr; b += NT) {
                        double acc = 0.0;
                        for (int s = 0; s < d; s++) {
                            double t = 0.0;
                            for (int a = 0; a < cl; a++) t += vec[a] * A[((size_t)s * cl + a) * cr + b];
                            acc += phi[s] * t;
                        }
                        vec2[b] = acc;
                    }
                    x_prev = x[j];
                    have_prev = true;
                } else {
                    // A_d[s][b] = sum_a v[a] A[s][a][b]
                    for (int e = tid; e < d * cr; e += NT) {
                        const int s = e / cr, b = e % cr;
                        double t = 0.0;
                        for (int a = 0; a < cl; a++) t += vec[a] * A[((size_t)s * cl + a) * cr + b];
                        Ad[s * ld + b] = t;
                    }
                    const double* Gk = GR + (size_t)k * P.chimax * P.chimax;
                    for (int e = tid; e < cr * cr; e += NT) Gm[(e / cr) * ld + (e % cr)] = Gk[e];
                    __syncthreads();
                    // T1[s][c] = sum_b Ad[s][b] G[b][c];  rho[s][t] = sum_c T1[s][c] Ad[t][c]
                    for (int e = tid; e < d * cr; e += NT) {
                        const int s = e / cr, c = e % cr;
                        double t = 0.0;
                        for (int b = 0; b < cr; b++) t += Ad[s * ld + b] * Gm[b * ld + c];
                        Td[s * ld + c] = t;
                    }
                    __syncthreads();
                    for (int e = tid; e < d * d; e += NT) {
                        const int s = e / d, t2 = e % d;
                        double t = 0.0;
                        for (int c = 0; c < cr; c++) t += Td[s * ld + c] * Ad[t2 * ld + c];
                        rho[e] = t;
                    }
                    __syncthreads();
                    // symmetrise + scale rho to unit max (p is defined up to a constant)
                    double mx = 0.0;
                    for (int e = tid; e < d * d; e += NT) mx = fmax(mx, fabs(rho[e]));
                    mx = block_reduce_max(mx, red);
                    const double rs = mx > 0.0 ? 1.0 / mx : 1.0;
                    for (int e = tid; e < d * d; e += NT) {
                        const int s = e / d, t2 = e % d;
                        if (s <= t2) { const double v = 0.5 * (rho[s * d + t2] + rho[t2 * d + s]) * rs; rho2[s * d + t2] = v; rho2[t2 * d + s] = v; }
                    }
                    __syncthreads();
                    for (int e = tid; e < d * d; e += NT) rho[e] = rho2[e];
                    __syncthreads();
                    // pdf on the grid: p[g] = || rho Phi_g ||^2   (thread-contiguous chunks for the scan)
                    const int per = (G + NT - 1) / NT;
                    const int g0 = tid * per, g1 = min(G, g0 + per);
                    for (int g = g0; g < g1; g++) {
                        const double* ph = P.genc + (size_t)g * d;
                        double pv = 0.0;
                        for (int s = 0; s < d; s++) {
                            double t = 0.0;
                            for (int t2 = 0; t2 < d; t2++) t += rho[s * d + t2] * ph[t2];
                            pv += t * t;
                        }
                        pbuf[g] = pv;
                    }
                    __syncthreads();
                    int gsel = 0;
                    double xsel = 0.0;
                    if (P.method == MPST_IMPUTE_MEDIAN || P.method == MPST_IMPUTE_ITS) {
                        // cumulative trapezoid c[g] = c[g-1] + (p[g-1] + p[g]), scaled by h = (x1-x0)/2
                        double loc = 0.0;
                        for (int g = max(g0, 1); g < g1; g++) loc += pbuf[g - 1] + pbuf[g];
                        // exclusive prefix over threads (NT partial sums, serial in thread 0: NT is small)
                        double* pre = scr;                                   // NT+1 doubles
                        pre[tid] = loc;
                        __syncthreads();
                        if (tid == 0) {
                            double run = 0.0;
                            for (int t2 = 0; t2 < NT; t2++) { const double v = pre[t2]; pre[t2] = run; run += v; }
                            pre[NT] = run;
                        }
                        __syncthreads();
                        const double h = (P.grid[1] - P.grid[0]) * 0.5;
                        const double Z = h * pre[NT];
                        const double u = (P.method == MPST_IMPUTE_MEDIAN) ? 0.5
                                         : P.uniforms[((size_t)inst * P.ntraj + tr) * P.Kmax + k];
                        double best = 1e300;
                        int bg = 0x7fffffff;
                        double run = pre[tid];
                        for (int g = g0; g < g1; g++) {
                            if (g > 0) run += pbuf[g - 1] + pbuf[g];
                            const double cg = h * run;
                            const double val = fabs(cg / Z - u);
                            if (val < best) { best = val; bg = g; }
                        }
                        // lexicographic (val, g) minimum over the block
                        double* bv = scr;
                        int* bi = reinterpret_cast<int*>(scr + NT + 2);
                        __syncthreads();
                        bv[tid] = best; bi[tid] = bg;
                        __syncthreads();
                        if (tid == 0) {
                            double b0 = bv[0]; int i0 = bi[0];
                            for (int t2 = 1; t2 < NT; t2++)
                                if (bv[t2] < b0 || (bv[t2] == b0 && bi[t2] < i0)) { b0 = bv[t2]; i0 = bi[t2]; }
                            s_misc[3] = i0;
                        }
                        __syncthreads();
                        gsel = s_misc[3];
                        if (gsel < 0 || gsel >= G) gsel = 0;          // non-finite pdf (NaN in the cores): stay in bounds
                        xsel = P.grid[gsel];
                    } else if (P.method == MPST_IMPUTE_MODE) {
                        double best = -1.0; int bg = 0x7fffffff;
                        double bestall = -1.0; int bgall = 0x7fffffff;
                        const bool filt = have_prev && P.max_jump >= 0.0;
                        for (int g = g0; g < g1; g++) {
                            const double pv = pbuf[g];
                            if (pv > bestall) { bestall = pv; bgall = g; }
                            if (!filt || fabs(P.grid[g] - x_prev) <= P.max_jump) if (pv > best) { best = pv; bg = g; }
                        }
                        double* bv = scr; int* bi = reinterpret_cast<int*>(scr + 2 * NT + 2);
                        __syncthreads();
                        bv[tid] = best; bi[tid] = bg; bv[NT + tid] = bestall; bi[NT + tid] = bgall;
                        __syncthreads();
                        if (tid == 0) {
                            double b0 = -1.0; int i0 = 0x7fffffff; double b1 = -1.0; int i1 = 0x7fffffff;
                            for (int t2 = 0; t2 < NT; t2++) {
                                if (bv[t2] > b0 || (bv[t2] == b0 && bi[t2] < i0)) { b0 = bv[t2]; i0 = bi[t2]; }
                                if (bv[NT + t2] > b1 || (bv[NT + t2] == b1 && bi[NT + t2] < i1)) { b1 = bv[NT + t2]; i1 = bi[NT + t2]; }
                            }
                            s_misc[3] = (i0 != 0x7fffffff) ? i0 : i1;      // no admissible point: global argmax (:137-141)
                        }
                        __syncthreads();
                        gsel = s_misc[3];
                        xsel = P.grid[gsel];
                    } else {   // mean (sampling_utils.jl:64-101): E[x] = sum x p dx / Z, Z = trapz
                        double sp = 0.0, sxp = 0.0;
                        for (int g = g0; g < g1; g++) {
                            const double pv = pbuf[g];
                            sxp += P.grid[g] * pv;
                            sp += (g == 0 || g == G - 1) ? 0.5 * pv : pv;
                        }
                        sp = block_reduce_sum(sp, red);
                        sxp = block_reduce_sum(sxp, red);
                        const double dx = (P.grid[G - 1] - P.grid[0]) / (double)(G - 1);     // mean(abs(diff(xvals)))
                        const double Z = (P.grid[1] - P.grid[0]) * sp;
                        xsel = sxp * dx / Z;
                        gsel = -1;
                    }
                    if (tid == 0) xo[j] = xsel;
                    // state of the chosen value, then v <- state . A_d
                    if (gsel >= 0) { for (int s = tid; s < d; s += NT) phi[s] = P.genc[(size_t)gsel * d + s]; }
                    else if (tid == 0) encode_any(P.basis, xsel, d, phi);
                    __syncthreads();
                    for (int b = tid; b < cr; b += NT) {
                        double t = 0.0;
                        for (int s = 0; s < d; s++) t += phi[s] * Ad[s * ld + b];
                        vec2[b] = t;
                    }
                    x_prev = xsel;
                    have_prev = true;
                    k++;
                }
                __syncthreads();
                double nn = 0.0;
                for (int b = tid; b < cr; b += NT) nn += vec2[b] * vec2[b];
                nn = block_reduce_sum(nn, red);
                const double sc = nn > 0.0 ? rsqrt(nn) : 1.0;
                for (int b = tid; b < P.chimax; b += NT) vec[b] = b < cr ? vec2[b] * sc : 0.0;
                __syncthreads();
            }
        }
        __syncthreads();
    }
}

// class slice of one core -> [s][a][b]
__global__ void slice_core_kernel(CoreView v, int d, int chi_l, int chi_r, int cls, double* __restrict__ dst) {
    const int64_t n = (int64_t)d * chi_l * chi_r;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int b = (int)(e % chi_r);
    const int a = (int)((e / chi_r) % chi_l);
    const int s = (int)(e / ((int64_t)chi_r * chi_l));
    dst[e] = v.p[s * v.ss + a * v.sa + b * v.sb + (int64_t)cls * v.sc];
}
}  // namespace

int impute_batch(mpst_ctx* c, int class_idx, const double* X, const uint8_t* missing, int64_t n, int method,
                 const double* xgrid, int G, const double* uniforms, int n_traj, double max_jump, double* out) {
    if (!X || !missing || !xgrid || !out || n < 0 || G < 2 || c->T == 0) { c->err = "impute_batch: bad arguments"; return MPST_E_INVALID; }
    if (class_idx < 0 || class_idx >= c->C) { c->err = "impute_batch: class index out of range"; return MPST_E_INVALID; }
    if (method < MPST_IMPUTE_MEDIAN || method > MPST_IMPUTE_ITS) { c->err = "impute_batch: unknown method"; return MPST_E_INVALID; }
    if (method == MPST_IMPUTE_ITS && !uniforms) { c->err = "impute_batch: ITS needs the uniform draws"; return MPST_E_INVALID; }
    if (c->have_phi || (c->basis != MPST_BASIS_LEGENDRE_NO_NORM && c->basis != MPST_BASIS_LEGENDRE_NORM && c->basis != MPST_BASIS_UNIFORM)) {
        c->err = "impute_batch: needs one of the on-device real bases";
        return MPST_E_UNSUPPORTED;
    }
    if (n == 0) return MPST_OK;
    if (method != MPST_IMPUTE_ITS) n_traj = 1;
    if (n_traj < 1) n_traj = 1;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int T = c->T, d = c->d;
    std::vector<int> chi(T + 1, 1);
    std::vector<int64_t> off(T, 0);
    int chimax = 1;
    int64_t tot = 0;
    for (int j = 0; j < T; j++) {
        const Core& k = c->cores[j];
        if (!k.dev) { c->err = "impute_batch: cores not set"; return MPST_E_INVALID; }
        chi[j] = k.chi_l; chi[j + 1] = k.chi_r;
        chimax = std::max(chimax, std::max(k.chi_l, k.chi_r));
        off[j] = tot;
        tot += (int64_t)d * k.chi_l * k.chi_r;
    }
    const int ld = chimax + 1;
    const size_t smem = sizeof(double) * ((size_t)4 * chimax * ld + 2 * chimax + 2 * (size_t)d * ld + 2 * (size_t)d * d + MPST_MAX_D + 32 +
                                          3 * NT + 16);
    if (smem > 227 * 1024) { c->err = "impute_batch: chi too large for the shared-memory Gram matrices (chi <= 76 at d = 16)"; return MPST_E_UNSUPPORTED; }
    // host-side Kmax
    int Kmax = 0;
    for (int64_t i = 0; i < n; i++) { int k = 0; for (int j = 0; j < T; j++) k += missing[i * T + j] ? 1 : 0; Kmax = std::max(Kmax, k); }
    if (Kmax == 0) { for (int64_t i = 0; i < n; i++) for (int tr = 0; tr < n_traj; tr++) memcpy(out + (i * n_traj + tr) * T, X + i * T, sizeof(double) * T); return MPST_OK; }
    const int grid = (int)std::min<int64_t>(n, 2 * (int64_t)c->sm_count);
    double *dcores = nullptr, *dX = nullptr, *dgrid = nullptr, *dgenc = nullptr, *dunif = nullptr, *dout = nullptr, *dGR = nullptr, *dp = nullptr;
    uint8_t* dmask = nullptr;
    int64_t* doff = nullptr;
    int* dchi = nullptr;
    int rc = MPST_OK;
    auto cleanup = [&]() {
        cudaFree(dcores); cudaFree(dX); cudaFree(dgrid); cudaFree(dgenc); cudaFree(dunif); cudaFree(dout); cudaFree(dGR); cudaFree(dp);
        cudaFree(dmask); cudaFree(doff); cudaFree(dchi);
    };
#define IMP_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { c->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cleanup(); return MPST_E_CUDA; } } while (0)
    IMP_TRY(cudaMalloc(&dcores, sizeof(double) * tot));
    IMP_TRY(cudaMalloc(&dX, sizeof(double) * n * T));
    IMP_TRY(cudaMalloc(&dmask, (size_t)n * T));
    IMP_TRY(cudaMalloc(&dgrid, sizeof(double) * G));
    IMP_TRY(cudaMalloc(&dgenc, sizeof(double) * (size_t)G * d));
    IMP_TRY(cudaMalloc(&dout, sizeof(double) * n * n_traj * T));
    IMP_TRY(cudaMalloc(&dGR, sizeof(double) * ((size_t)grid * Kmax * chimax * chimax + 64)));
    IMP_TRY(cudaMalloc(&dp, sizeof(double) * (size_t)grid * G));
    IMP_TRY(cudaMalloc(&doff, sizeof(int64_t) * T));
    IMP_TRY(cudaMalloc(&dchi, sizeof(int) * (T + 1)));
    if (uniforms) {
        IMP_TRY(cudaMalloc(&dunif, sizeof(double) * n * n_traj * Kmax));
        IMP_TRY(cudaMemcpyAsync(dunif, uniforms, sizeof(double) * n * n_traj * Kmax, cudaMemcpyHostToDevice, c->stream));
    }
    IMP_TRY(cudaMemcpyAsync(dX, X, sizeof(double) * n * T, cudaMemcpyHostToDevice, c->stream));
    IMP_TRY(cudaMemcpyAsync(dmask, missing, (size_t)n * T, cudaMemcpyHostToDevice, c->stream));
    IMP_TRY(cudaMemcpyAsync(dgrid, xgrid, sizeof(double) * G, cudaMemcpyHostToDevice, c->stream));
    IMP_TRY(cudaMemcpyAsync(doff, off.data(), sizeof(int64_t) * T, cudaMemcpyHostToDevice, c->stream));
    IMP_TRY(cudaMemcpyAsync(dchi, chi.data(), sizeof(int) * (T + 1), cudaMemcpyHostToDevice, c->stream));
    for (int j = 0; j < T; j++) {
        const Core& k = c->cores[j];
        CoreView v;
        v.p = k.dev; v.ss = 1;
        if (k.orient == ORIENT_LEFT) { v.sa = d; v.sb = (long)d * k.chi_l; } else { v.sb = d; v.sa = (long)d * k.chi_r; }
        v.sc = k.has_label ? (long)d * k.chi_l * k.chi_r : 0;
        const int64_t ne = (int64_t)d * k.chi_l * k.chi_r;
        slice_core_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, c->stream>>>(v, d, k.chi_l, k.chi_r, k.has_label ? class_idx : 0, dcores + off[j]);
        c->launches++;
    }
    rc = launch_encode(c, c->basis, d, dgrid, G, dgenc, d);       // grid states, imputation.jl:92-107
    if (rc != MPST_OK) { cleanup(); return rc; }
    ImpParams P;
    P.cores = dcores; P.core_off = doff; P.chi = dchi; P.X = dX; P.mask = dmask; P.grid = dgrid; P.genc = dgenc;
    P.uniforms = dunif; P.out = dout; P.gr_scratch = dGR; P.p_scratch = dp;
    P.T = T; P.d = d; P.G = G; P.ntraj = n_traj; P.Kmax = Kmax; P.chimax = chimax; P.method = method; P.basis = c->basis;
    P.n = n; P.max_jump = max_jump;
    IMP_TRY(cudaFuncSetAttribute(impute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(c, MPST_T_IMPUTE);
    impute_kernel<<<grid, NT, smem, c->stream>>>(P);
    prof_end(c, MPST_T_IMPUTE);
    c->launches++;
    IMP_TRY(cudaGetLastError());
    IMP_TRY(cudaMemcpyAsync(out, dout, sizeof(double) * n * n_traj * T, cudaMemcpyDeviceToHost, c->stream));
    IMP_TRY(cudaStreamSynchronize(c->stream));
#undef IMP_TRY
    cleanup();
    return MPST_OK;
}
