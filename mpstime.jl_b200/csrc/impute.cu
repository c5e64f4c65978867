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
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cstdio>
#include "mpst_common.cuh"
#include "dmma.cuh"
#include "encode_device.cuh"
#include "encode_table_device.cuh"

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
    int T, d, G, ntraj, Kmax, chimax, method, basis, debug;
    int dbuf;                       // 1: a second slice buffer fits in shared memory (asynchronous double-buffered staging)
    int64_t n;
    double max_jump;
    double* err;                    // [n][ntraj][T] error bars (WMAD / std) or null
    int get_err;                    // median: get_wmad, mean: get_std
    int max_trials;                 // ITS with rejection: draws per site
    double rej_thr;                 // ITS: accept when |x - median| < rej_thr * WMAD; < 0 = :none
    int64_t ustride;                // rejection mode: uniforms per instance (flat stream, consumed in order)
    // series pdf: nq = 2d-1 Chebyshev-Gauss nodes; leg_pq[l][i] = P_l(node_i) (d x nq), leg_wq[k][i] =
    // (k == 0 ? 1 : 2) / nq T_k(node_i) (nq x nq, the Chebyshev analysis matrix); null = direct evaluation
    const double* leg_pq;
    const double* leg_wq;
    int nq;
    // table encodings (data-driven / time-dependent bases, encode_table.cu): genc holds one [G][d] block of grid states
    // per chain position (genc_stride = G*d, or 0 when every site shares one table / for the built-in bases); known and
    // imputed values are encoded on the fly with the tables of their site (site = position, mirrored when `mirror`)
    int64_t genc_stride;
    int tab_kind, tab_nsites, mirror;
    int64_t tab_istride, tab_dstride;
    const int* tab_ip;
    const double* tab_dp;
};

__device__ __forceinline__ void encode_any(int basis, double x, int d, double* v) {
    if (basis == MPST_BASIS_LEGENDRE_NO_NORM) encode_point<MPST_BASIS_LEGENDRE_NO_NORM>(x, d, v);
    else if (basis == MPST_BASIS_LEGENDRE_NORM) encode_point<MPST_BASIS_LEGENDRE_NORM>(x, d, v);
    else encode_point<MPST_BASIS_UNIFORM>(x, d, v);
}
// state of value x at chain position j (TAB: a table encoding; the built-in instantiation carries none of that code)
template <bool TAB>
__device__ __forceinline__ void state_at(const ImpParams& P, int j, double x, double* v) {
    if (TAB) {
        const int site = P.tab_nsites == 1 ? 0 : (P.mirror ? P.T - 1 - j : j);
        table_point(P.tab_kind, x, P.d, P.tab_ip + (size_t)site * P.tab_istride, P.tab_dp + (size_t)site * P.tab_dstride, v);
    } else encode_any(P.basis, x, P.d, v);
}

// C[n8 x n8] (+)= A * B  (or A * B^T when TB) on the FP64 tensor cores; all operands in shared memory with pitch
// ld == 4 (mod 16) doubles (conflict-free DMMA fragment loads), dimensions padded with zeros to a multiple of 8.
// 8 warps as 4 (block-row pairs) x 2 (block-column quads).
template <bool TB, bool ACC>
__device__ __forceinline__ void smem_dmma_matmul(double* __restrict__ Cm, const double* __restrict__ A,
                                                 const double* __restrict__ B, int n8, int ld) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int nb = n8 >> 3;
    // warp (wr, wc) owns 2 block rows x 4 block columns at a time: per k-step 2 A + 4 B fragment loads feed 8 DMMAs
    // (the 1 x 8 blocking this replaces needed 9 loads for 8 DMMAs and made every warp stream all of B)
    const int wr = warp >> 1, wc = warp & 1;
    for (int bi = 2 * wr; bi < nb; bi += 8) {
        const bool r1 = bi + 1 < nb;
        for (int bj = 4 * wc; bj < nb; bj += 8) {
            double c[2][4][2];
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int i = (bi + u) * 8 + fr, j = (bj + t) * 8 + 2 * fc;
                    if (ACC && bi + u < nb && bj + t < nb) { c[u][t][0] = Cm[i * ld + j]; c[u][t][1] = Cm[i * ld + j + 1]; }
                    else { c[u][t][0] = 0.0; c[u][t][1] = 0.0; }
                }
            const double* a0p = A + (bi * 8 + fr) * ld + fc;
            const double* a1p = A + ((r1 ? bi + 1 : bi) * 8 + fr) * ld + fc;
#pragma unroll 2
            for (int k0 = 0; k0 < n8; k0 += 4) {
                const double av0 = a0p[k0], av1 = a1p[k0];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (bj + t < nb) {
                        const int jb = (bj + t) * 8;
                        const double bv = TB ? B[(jb + fr) * ld + k0 + fc] : B[(k0 + fc) * ld + jb + fr];
                        dmma_8x8x4(c[0][t][0], c[0][t][1], av0, bv);
                        dmma_8x8x4(c[1][t][0], c[1][t][1], av1, bv);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (bi + u < nb && bj + t < nb) {
                        const int i = (bi + u) * 8 + fr, j = (bj + t) * 8 + 2 * fc;
                        Cm[i * ld + j] = c[u][t][0];
                        Cm[i * ld + j + 1] = c[u][t][1];
                    }
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

// Bonnet recurrence P_l = a_l x P_{l-1} - b_l P_{l-2},  a_l = (2l-1)/l,  b_l = (l-1)/l
__constant__ double c_leg_a[MPST_MAX_D];
__constant__ double c_leg_b[MPST_MAX_D];

// p[g] = || rho Phi(x_g) ||^2 for the Legendre bases, NP grid points per thread at a time.  R[s][l] = rho[s][l] *
// sqrt((2l+1)/2) (zero padded to D x D) sits in shared memory (broadcast reads); P_l(x) comes from the Bonnet
// recurrence with compile-time coefficients, accumulators stay in registers.
template <int D, int NP>
__device__ __forceinline__ void pdf_legendre(const double* __restrict__ Rt, const double* __restrict__ grid, int g0, int g1,
                                             double* __restrict__ pbuf) {
    // Rt[l][s] (s contiguous, 16-byte aligned rows): one LDS.128 feeds 2*NP DFMAs
    for (int g = g0; g < g1; g += NP) {
        double x[NP], pm[NP], pc[NP], t[NP][D];
#pragma unroll
        for (int q = 0; q < NP; q++) {
            x[q] = grid[min(g + q, g1 - 1)];
            pm[q] = 1.0; pc[q] = 1.0;
#pragma unroll
            for (int s = 0; s < D; s++) t[q][s] = 0.0;
        }
#pragma unroll 1
        for (int l = 0; l < D; l++) {
            if (l == 1) {
#pragma unroll
                for (int q = 0; q < NP; q++) pc[q] = x[q];
            } else if (l > 1) {
                const double al = c_leg_a[l], bl = c_leg_b[l];
#pragma unroll
                for (int q = 0; q < NP; q++) { const double pn = al * x[q] * pc[q] - bl * pm[q]; pm[q] = pc[q]; pc[q] = pn; }
            }
#pragma unroll
            for (int s = 0; s < D; s += 2) {
                const double2 r = *reinterpret_cast<const double2*>(Rt + l * D + s);
#pragma unroll
                for (int q = 0; q < NP; q++) { t[q][s] += r.x * pc[q]; t[q][s + 1] += r.y * pc[q]; }
            }
        }
#pragma unroll
        for (int q = 0; q < NP; q++) {
            if (g + q < g1) {
                double pv = 0.0;
#pragma unroll
                for (int s = 0; s < D; s++) pv += t[q][s] * t[q][s];
                pbuf[g + q] = pv;
            }
        }
    }
}

// p[g] = sum_k q_k T_k(x_g): the pdf || rho Phi(x) ||^2 of a Legendre basis of dimension d is a polynomial of degree
// 2(d-1) in x, so after an exact change to its own CHEBYSHEV series (2d-1 coefficients, see the caller) a grid point
// costs 2(2d-1) flop-instructions (Clenshaw: b_k = q_k + 2x b_{k+1} - b_{k+2}, one add and one FMA per term; the
// Legendre recurrence needs three) instead of ~d^2 + 2d.  NP points per thread at a time for ILP.
template <int NP>
__device__ __forceinline__ void pdf_legendre_series(const double* __restrict__ q, int nq, const double* __restrict__ grid,
                                                    int g0, int g1, double* __restrict__ pbuf) {
    for (int g = g0; g < g1; g += NP) {
        double x[NP], y[NP], b1[NP], b2[NP];
#pragma unroll
        for (int u = 0; u < NP; u++) { x[u] = grid[min(g + u, g1 - 1)]; y[u] = 2.0 * x[u]; b1[u] = 0.0; b2[u] = 0.0; }
#pragma unroll 1
        for (int k = nq - 1; k >= 1; k--) {
            const double qk = q[k];
#pragma unroll
            for (int u = 0; u < NP; u++) {
                const double t = fma(y[u], b1[u], qk - b2[u]);
                b2[u] = b1[u];
                b1[u] = t;
            }
        }
        const double q0 = q[0];
#pragma unroll
        for (int u = 0; u < NP; u++)
            if (g + u < g1) pbuf[g + u] = fmax(fma(x[u], b1[u], q0 - b2[u]), 0.0);
    }
}

// exclusive prefix sum of one value per thread over the block (warp shuffles + one smem hop); total in `total`
__device__ __forceinline__ double block_excl_scan(double v, double* scr, double& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) scr[warp] = inc;
    __syncthreads();
    double off = 0.0, tot = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) {
        const double t = scr[w];
        if (w < warp) off += t;
        tot += t;
    }
    total = tot;
    return off + inc - v;
}

// block-wide lexicographic best of (value, index): smallest (MAXI = false) or largest (MAXI = true) value, ties to the
// smaller index.  Every thread gets the winning index.
template <bool MAXI>
__device__ __forceinline__ int block_argbest(double v, int i, double* scr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const int oi = __shfl_down_sync(0xffffffffu, i, o);
        const bool better = MAXI ? (ov > v) : (ov < v);
        if (better || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    int* si = reinterpret_cast<int*>(scr + NT / 32);
    __syncthreads();
    if (lane == 0) { scr[warp] = v; si[warp] = i; }
    __syncthreads();
    v = scr[0]; i = si[0];
#pragma unroll
    for (int w = 1; w < NT / 32; w++) {
        const double ov = scr[w];
        const int oi = si[w];
        const bool better = MAXI ? (ov > v) : (ov < v);
        if (better || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    return i;
}

// stage slice s of a core ([s][a][b], cl x cr) into the zero-padded n8 x n8 shared tile
__device__ __forceinline__ void stage_slice(double* __restrict__ As, const double* __restrict__ A, int s, int cl, int cr,
                                            int n8, int ld) {
    const double* src = A + (size_t)s * cl * cr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int a = warp; a < n8; a += NT / 32)
        for (int b = lane; b < n8; b += 32) As[a * ld + b] = (a < cl && b < cr) ? src[a * cr + b] : 0.0;
}

// asynchronous variant: 16-byte cp.async chunks of the valid cl x cr region (cr even, 16-byte aligned rows); the zero
// padding of the buffer is written once per site by the caller
__device__ __forceinline__ void stage_slice_async(double* __restrict__ As, const double* __restrict__ A, int s, int cl, int cr,
                                                  int ld) {
    const double* src = A + (size_t)s * cl * cr;
    const int half = cr >> 1;                          // 16-byte chunks per row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int a = warp; a < cl; a += NT / 32)           // a row per warp at a time: no index division
        for (int b2 = lane; b2 < half; b2 += 32) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(As + a * ld + 2 * b2);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)a * cr + 2 * b2) : "memory");
        }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool TAB>
__global__ void __launch_bounds__(NT, 1) impute_kernel(ImpParams P) {
    extern __shared__ double sm[];
    const int n8 = (P.chimax + 7) & ~7;
    const int ld = n8 + 4;                             // == 4 (mod 16) when n8 is a multiple of 16; see launch check
    const int msz = n8 * ld;
    double* Gm = sm;                                   // current right Gram
    double* Gn = Gm + msz;                             // next / accumulator
    double* As = Gn + msz;                             // staged A_s / M
    double* T1 = As + msz;                             // temp product
    double* vec = T1 + msz;                            // [n8] left / right vector
    double* vec2 = vec + n8;
    double* Ad = vec2 + n8;                            // [d][ld]
    double* Td = Ad + P.d * ld;                        // [d][ld]
    double* rho = Td + P.d * ld;                       // [d][d]
    double* rho2 = rho + P.d * P.d;                    // [d][d]
    double* phi = rho2 + P.d * P.d;                    // [d]
    double* red = phi + MPST_MAX_D;                    // [32] reductions
    double* scr = red + 32;                            // [max(3*NT + 16, 4*n8)] scan / partial-sum scratch
    double* pdfR = scr + max(3 * NT + 16, 4 * n8);     // [32*32] rho * normalisation for the Legendre pdf
    double* As2 = pdfR + MPST_MAX_D * MPST_MAX_D;      // second slice buffer (only when P.dbuf)
    __shared__ int s_misc[4];
    const int tid = threadIdx.x;
    const int T = P.T, d = P.d, G = P.G;
    const int grp = tid >> 6, l64 = tid & 63;          // 4 groups of 64 threads for the vector contractions
    double* GR = P.gr_scratch + (size_t)blockIdx.x * P.Kmax * P.chimax * P.chimax;
    double* pbuf = P.p_scratch + (size_t)blockIdx.x * G;

    // acc[q] += wgt * sum_a vec[a] * src[a*cr + j],  j = l64 + 64 q, straight from global memory (whole contraction
    // index per thread: no cross-group reduction)
    auto colsum_global = [&](const double* __restrict__ src, int cl, int cr, double wgt, double* acc) {
        for (int q = 0, jj = l64; jj < cr; jj += 64, q++) {
            double t = 0.0;
#pragma unroll 8
            for (int a = 0; a < cl; a++) t = fma(vec[a], __ldg(src + (size_t)a * cr + jj), t);
            acc[q] += wgt * t;
        }
    };
    auto reduce_partials = [&](const double* acc, int n, double* out) {     // out[i] = sum over the 4 groups
        for (int q = 0, i = l64; i < n8; i += 64, q++) scr[grp * n8 + i] = (i < n) ? acc[q] : 0.0;
        __syncthreads();
        for (int i = tid; i < n8; i += NT) out[i] = scr[i] + scr[n8 + i] + scr[2 * n8 + i] + scr[3 * n8 + i];
        __syncthreads();
    };
    auto normalise_into_vec = [&](int n) {                                   // vec <- vec2 / ||vec2||
        double nn = 0.0;
        for (int i = tid; i < n; i += NT) nn += vec2[i] * vec2[i];
        nn = block_reduce_sum(nn, red);
        const double sc = nn > 0.0 ? rsqrt(nn) : 1.0;
        for (int i = tid; i < n8; i += NT) vec[i] = i < n ? vec2[i] * sc : 0.0;
        __syncthreads();
    };

    long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tq = 0;
#define TICK() do { if (P.debug) tq = clock64(); } while (0)
#define TOCK(i) do { if (P.debug) tm[i] += clock64() - tq; } while (0)
    for (int64_t inst = blockIdx.x; inst < P.n; inst += gridDim.x) {
        const double* x = P.X + inst * T;
        const uint8_t* mk = P.mask + inst * T;
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
        for (int b = tid; b < n8; b += NT) vec[b] = (b == 0) ? 1.0 : 0.0;
        __syncthreads();
        TICK();
        for (int j = T - 1; j > last; j--) {
            const int cl = P.chi[j], cr = P.chi[j + 1];
            if (tid == 0) state_at<TAB>(P, j, x[j], phi);
            const double* A = P.cores + P.core_off[j];           // [s][a][b]
            __syncthreads();
            // r'[a] = sum_s phi_s sum_b A_s[a][b] r[b]: one row a per warp at a time, lanes along b (coalesced reads
            // straight from global memory -- every core element is used once), one shuffle reduction per row
            for (int a = tid >> 5; a < cl; a += NT / 32) {
                double t = 0.0;
                const double* row0 = A + (size_t)a * cr;
                const size_t sstride = (size_t)cl * cr;
                for (int b = tid & 31; b < cr; b += 32) {
                    const double vb = vec[b];
                    double tb = 0.0;
#pragma unroll 8
                    for (int s = 0; s < d; s++) tb = fma(phi[s], __ldg(row0 + s * sstride + b), tb);   // d loads in flight
                    t = fma(tb, vb, t);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if ((tid & 31) == 0) vec2[a] = t;
            }
            __syncthreads();
            normalise_into_vec(cl);
        }
        TOCK(0);
        TICK();
        {   // G = r r^T  (dimension chi[last+1]), zero padded
            const int cr = P.chi[last + 1];
            for (int e = tid; e < n8 * n8; e += NT) {
                const int a = e / n8, b = e - a * n8;
                Gm[a * ld + b] = (a < cr && b < cr) ? vec[a] * vec[b] : 0.0;
            }
            __syncthreads();
        }
        int kidx = K - 1;
        for (int j = last; j >= first; j--) {
            const int cl = P.chi[j], cr = P.chi[j + 1];
            const double* A = P.cores + P.core_off[j];
            if (mk[j]) {
                double* dst = GR + (size_t)kidx * P.chimax * P.chimax;     // the Gram this missing site sees
                for (int e = tid; e < cr * cr; e += NT) dst[e] = Gm[(e / cr) * ld + (e % cr)];
                kidx--;
                if (j == first) break;
                for (int e = tid; e < n8 * n8; e += NT) Gn[(e / n8) * ld + (e % n8)] = 0.0;
                if (P.dbuf && !(cr & 1) && !((cl * cr) & 1) && !(P.core_off[j] & 1)) {
                    // double-buffered: slice s+1 streams in (cp.async) while the two products of slice s run
                    __syncthreads();
                    if (cl < n8 || cr < n8)
                        for (int e = tid; e < n8 * n8; e += NT) { const int o = (e / n8) * ld + (e % n8); As[o] = 0.0; As2[o] = 0.0; }
                    __syncthreads();
                    stage_slice_async(As, A, 0, cl, cr, ld);
                    for (int s = 0; s < d; s++) {
                        double* cur = (s & 1) ? As2 : As;
                        stage_wait_all();
                        __syncthreads();                                       // slice s visible, products of slice s-1 done
                        if (s + 1 < d) stage_slice_async((s & 1) ? As : As2, A, s + 1, cl, cr, ld);
                        smem_dmma_matmul<false, false>(T1, cur, Gm, n8, ld);   // T1 = A_s G
                        __syncthreads();
                        smem_dmma_matmul<true, true>(Gn, T1, cur, n8, ld);     // Gn += T1 A_s^T
                    }
                } else {
                    for (int s = 0; s < d; s++) {
                        __syncthreads();
                        stage_slice(As, A, s, cl, cr, n8, ld);
                        __syncthreads();
                        smem_dmma_matmul<false, false>(T1, As, Gm, n8, ld);    // T1 = A_s G
                        __syncthreads();
                        smem_dmma_matmul<true, true>(Gn, T1, As, n8, ld);      // Gn += T1 A_s^T
                    }
                }
            } else {
                if (tid == 0) state_at<TAB>(P, j, x[j], phi);
                for (int e = tid; e < n8 * n8; e += NT) Gn[(e / n8) * ld + (e % n8)] = 0.0;   // M accumulates in Gn
                for (int s = 0; s < d; s++) {
                    __syncthreads();
                    stage_slice(As, A, s, cl, cr, n8, ld);
                    __syncthreads();
                    const double ps = phi[s];
                    for (int e = tid; e < n8 * n8; e += NT) { const int o = (e / n8) * ld + (e % n8); Gn[o] += ps * As[o]; }
                }
                __syncthreads();
                for (int e = tid; e < n8 * n8; e += NT) { const int o = (e / n8) * ld + (e % n8); As[o] = Gn[o]; }
                __syncthreads();
                smem_dmma_matmul<false, false>(T1, As, Gm, n8, ld);            // T1 = M G
                __syncthreads();
                smem_dmma_matmul<true, false>(Gn, T1, As, n8, ld);             // Gn = T1 M^T
            }
            __syncthreads();
            // renormalise (scale invariance) and swap
            double mx = 0.0;
            for (int e = tid; e < n8 * n8; e += NT) mx = fmax(mx, fabs(Gn[(e / n8) * ld + (e % n8)]));
            mx = block_reduce_max(mx, red);
            const double sc = mx > 0.0 ? 1.0 / mx : 1.0;
            for (int e = tid; e < n8 * n8; e += NT) { const int o = (e / n8) * ld + (e % n8); Gm[o] = Gn[o] * sc; }
            __syncthreads();
        }
        __syncthreads();
        TOCK(1);

        // ---------------- forward pass(es) ---------------------------------------------------------
        int64_t ucur = 0;                                     // position in this instance's uniform stream (rejection mode)
        for (int tr = 0; tr < P.ntraj; tr++) {
            double* xo = P.out + (inst * P.ntraj + tr) * T;
            for (int a = tid; a < n8; a += NT) vec[a] = (a == 0) ? 1.0 : 0.0;
            __syncthreads();
            int k = 0;
            // x_prev of impute_at! (MPS_methods.jl:136-144, 158): the value next to the first site to impute if there
            // is one, afterwards always the value imputed last -- known sites in between do NOT update it
            double x_prev = first > 0 ? x[first - 1] : 0.0;
            bool have_prev = first > 0;
            for (int j = 0; j <= last; j++) {
                const int cl = P.chi[j], cr = P.chi[j + 1];
                const double* A = P.cores + P.core_off[j];
                TICK();
                if (!mk[j]) {
                    if (tid == 0) state_at<TAB>(P, j, x[j], phi);
                    __syncthreads();
                    // v'[b] = sum_s phi_s sum_a v[a] A_s[a][b]: every core element is used once, so it is read straight
                    // from global memory (coalesced over b, loads batched); the 4 thread groups take the slices s = g mod 4
                    double acc[2] = {0.0, 0.0};
                    for (int s = grp; s < d; s += 4) colsum_global(A + (size_t)s * cl * cr, cl, cr, phi[s], acc);
                    reduce_partials(acc, cr, vec2);
                    TOCK(2);
                } else {
                    // A_d[s][b] = sum_a v[a] A[s][a][b]
                    for (int s = grp; s < d; s += 4) {                // one slice per thread group, no reduction needed
                        double acc[2] = {0.0, 0.0};
                        colsum_global(A + (size_t)s * cl * cr, cl, cr, 1.0, acc);
                        for (int q = 0, jj = l64; jj < cr; jj += 64, q++) Ad[s * ld + jj] = acc[q];
                    }
                    __syncthreads();
                    TOCK(3);
                    TICK();
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
                    TOCK(4);
                    TICK();
                    // pdf on the grid: p[g] = || rho Phi_g ||^2   (thread-contiguous chunks for the scan)
                    const int per = (G + NT - 1) / NT;
                    const int g0 = tid * per, g1 = min(G, g0 + per);
                    if (!TAB && (P.basis == MPST_BASIS_LEGENDRE_NO_NORM || P.basis == MPST_BASIS_LEGENDRE_NORM)) {
                        // R = rho * diag(sqrt((2l+1)/2)), zero padded to the next supported order (in scr/rho2 space)
                        const int D = d <= 8 ? 8 : d <= 16 ? 16 : d <= 24 ? 24 : 32;
                        double* Rm = pdfR;
                        for (int e = tid; e < D * D; e += NT) {
                            const int l = e / D, s = e - l * D;
                            Rm[e] = (s < d && l < d) ? rho[s * d + l] * sqrt((double)(2 * l + 1) * 0.5) : 0.0;
                        }
                        __syncthreads();
                        if (P.leg_pq != nullptr) {
                            // exact Chebyshev series of the degree-2(d-1) polynomial p: sample it at the nq = 2d-1
                            // Chebyshev-Gauss nodes (d^2 flops each, nq points instead of G), discrete cosine transform
                            // (exact for degree <= nq - 1), then Clenshaw on the grid
                            const int nq = P.nq;
                            for (int i = tid; i < nq; i += NT) {
                                double pv = 0.0;
                                for (int s2 = 0; s2 < d; s2++) {
                                    double t = 0.0;
                                    for (int l = 0; l < d; l++) t = fma(Rm[l * D + s2], P.leg_pq[l * nq + i], t);
                                    pv = fma(t, t, pv);
                                }
                                scr[i] = pv;
                            }
                            __syncthreads();
                            for (int k2 = tid; k2 < nq; k2 += NT) {
                                double qv = 0.0;
                                for (int i = 0; i < nq; i++) qv = fma(P.leg_wq[k2 * nq + i], scr[i], qv);
                                scr[64 + k2] = qv;
                            }
                            __syncthreads();
                            pdf_legendre_series<4>(scr + 64, nq, P.grid, g0, g1, pbuf);
                        }
                        else if (D == 8) pdf_legendre<8, 4>(Rm, P.grid, g0, g1, pbuf);
                        else if (D == 16) pdf_legendre<16, 4>(Rm, P.grid, g0, g1, pbuf);
                        else if (D == 24) pdf_legendre<24, 2>(Rm, P.grid, g0, g1, pbuf);
                        else pdf_legendre<32, 2>(Rm, P.grid, g0, g1, pbuf);
                    } else {
                        for (int g = g0; g < g1; g++) {
                            double ph[MPST_MAX_D];
                            if (TAB) { for (int t2 = 0; t2 < d; t2++) ph[t2] = P.genc[(size_t)j * P.genc_stride + (size_t)g * d + t2]; }
                            else encode_any(P.basis, P.grid[g], d, ph);
                            double pv = 0.0;
                            for (int s = 0; s < d; s++) {
                                double t = 0.0;
                                for (int t2 = 0; t2 < d; t2++) t += rho[s * d + t2] * ph[t2];
                                pv += t * t;
                            }
                            pbuf[g] = pv;
                        }
                    }
                    __syncthreads();
                    TOCK(5);
                    TICK();
                    int gsel = 0;
                    double xsel = 0.0;
                    double errv = 0.0;
                    if (P.method == MPST_IMPUTE_MEDIAN || P.method == MPST_IMPUTE_ITS) {
                        // cumulative trapezoid c[g] = c[g-1] + (p[g-1] + p[g]), scaled by h = (x1-x0)/2
                        double loc = 0.0;
                        {
                            double prev = (g0 > 0 && g0 < g1) ? pbuf[g0 - 1] : 0.0;          // every value is read once
                            for (int g = g0; g < g1; g++) {
                                const double cur = pbuf[g];
                                if (g > 0) loc += prev + cur;
                                prev = cur;
                            }
                        }
                        double tot;
                        const double pre_t = block_excl_scan(loc, scr, tot);
                        const double h = (P.grid[1] - P.grid[0]) * 0.5;
                        const double Z = h * tot;
                        // grid point whose cdf is closest to u: argmin |cdf/Z - u| without a division per point
                        auto pick = [&](double u) -> int {
                            double best = 1e300;
                            int bg = 0x7fffffff;
                            double run = pre_t;
                            const double uZ = u * Z;
                            double prev = (g0 > 0 && g0 < g1) ? pbuf[g0 - 1] : 0.0;
                            for (int g = g0; g < g1; g++) {
                                const double cur = pbuf[g];
                                if (g > 0) run += prev + cur;
                                prev = cur;
                                const double val = fabs(h * run - uZ);
                                if (val < best) { best = val; bg = g; }
                            }
                            int gs = block_argbest<false>(best, bg, scr);
                            if (gs < 0 || gs >= G) gs = 0;            // non-finite pdf (NaN in the cores): stay in bounds
                            return gs;
                        };
                        // weighted median absolute deviation about grid point m (sampling_utils.jl:192-196 ->
                        // StatsBase median(v, pweights(w)) = quantile(v, w, 0.5)): the (|x_g - x_m|, p_g/Z) pairs sorted by
                        // value then weight are m itself followed by the two points at grid distance 1, 2, ...; cumulative
                        // weight S_k, h = (sum w - w_1)/2 + w_1, linear interpolation at the first S_k > h
                        auto wmad = [&](int m) -> double {
                            const double xm = P.grid[m];
                            const double iZ = 1.0 / Z;
                            const int J = max(m, G - 1 - m) + 1;      // distances 0 .. J-1
                            const int perj = (J + NT - 1) / NT;
                            const int j0 = min(J, tid * perj), j1 = min(J, j0 + perj);
                            auto pairw = [&](int j, double* v, double* w) -> int {   // elements at distance j in sorted order
                                if (j == 0) { v[0] = 0.0; w[0] = pbuf[m] * iZ; return 1; }
                                int cnt = 0;
                                if (m - j >= 0) { v[cnt] = fabs(P.grid[m - j] - xm); w[cnt] = pbuf[m - j] * iZ; cnt++; }
                                if (m + j < G) { v[cnt] = fabs(P.grid[m + j] - xm); w[cnt] = pbuf[m + j] * iZ; cnt++; }
                                if (cnt == 2 && (v[1] < v[0] || (v[1] == v[0] && w[1] < w[0]))) {
                                    const double tv = v[0], tw = w[0]; v[0] = v[1]; w[0] = w[1]; v[1] = tv; w[1] = tw;
                                }
                                return cnt;
                            };
                            double lsum = 0.0;
                            for (int j = j0; j < j1; j++) { double v[2], w[2]; const int cnt = pairw(j, v, w); for (int q = 0; q < cnt; q++) lsum += w[q]; }
                            double wsum;
                            double S = block_excl_scan(lsum, scr, wsum);
                            const double w1 = pbuf[m] * iZ;
                            const double hq = 0.5 * (wsum - w1) + w1;
                            double vprev = 0.0;
                            if (j0 > 0 && j0 < J) { double v[2], w[2]; const int cnt = pairw(j0 - 1, v, w); vprev = v[cnt - 1]; }
                            int found = 0x7fffffff;
                            double res = 0.0;
                            for (int j = j0; j < j1 && found == 0x7fffffff; j++) {
                                double v[2], w[2];
                                const int cnt = pairw(j, v, w);
                                for (int q = 0; q < cnt; q++) {
                                    if (w[q] == 0.0) continue;                       // zero weights are dropped before sorting
                                    const double Sold = S;
                                    S += w[q];
                                    if (S > hq) { found = 2 * j + q; res = vprev + (hq - Sold) / (S - Sold) * (v[q] - vprev); break; }
                                    vprev = v[q];
                                }
                            }
                            const int win = block_argbest<false>((double)found, tid, scr);      // thread holding the first crossing
                            __syncthreads();
                            if (tid == win) scr[0] = (found == 0x7fffffff) ? fmax(fabs(P.grid[0] - xm), fabs(P.grid[G - 1] - xm)) : res;
                            __syncthreads();
                            const double out = scr[0];
                            __syncthreads();
                            return out;
                        };
                        if (P.method == MPST_IMPUTE_MEDIAN) {
                            gsel = pick(0.5);
                            if (P.get_err) errv = wmad(gsel);
                        } else if (P.rej_thr < 0.0) {
                            gsel = pick(P.uniforms[((size_t)inst * P.ntraj + tr) * P.Kmax + k]);
                        } else {
                            // rejection sampling about the median (sampling_utils.jl:291-311): the flat uniform stream of
                            // this instance is consumed in order, over sites and trajectories
                            const int gm = pick(0.5);
                            const double wm = wmad(gm);
                            gsel = gm;
                            for (int trial = 0; trial < P.max_trials; trial++) {
                                const double u = P.uniforms[(size_t)inst * P.ustride + ucur];
                                ucur++;
                                gsel = pick(u);
                                if (fabs(P.grid[gsel] - P.grid[gm]) < P.rej_thr * wm) break;
                            }
                            errv = wm;
                        }
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
                        const int i0 = block_argbest<true>(best, bg, scr);
                        const int i1 = block_argbest<true>(bestall, bgall, scr);
                        gsel = (i0 != 0x7fffffff) ? i0 : i1;               // no admissible point: global argmax (:137-141)
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
                        if (P.get_err) {                                  // sampling_utils.jl:89-97: sqrt(sum (x - E)^2 p dx / Z)
                            double sv = 0.0;
                            for (int g = g0; g < g1; g++) { const double df = P.grid[g] - xsel; sv += df * df * pbuf[g]; }
                            sv = block_reduce_sum(sv, red);
                            errv = sqrt(sv * dx / Z);
                        }
                    }
                    TOCK(6);
                    if (tid == 0) {
                        xo[j] = xsel;
                        if (P.err) P.err[(inst * P.ntraj + tr) * T + j] = errv;
                    }
                    // state of the chosen value, then v <- state . A_d
                    if (gsel >= 0) { for (int s = tid; s < d; s += NT) phi[s] = P.genc[(TAB ? (size_t)j * P.genc_stride : (size_t)0) + (size_t)gsel * d + s]; }
                    else if (tid == 0) state_at<TAB>(P, j, xsel, phi);
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
                for (int b = tid; b < n8; b += NT) vec[b] = b < cr ? vec2[b] * sc : 0.0;
                __syncthreads();
            }
        }
        __syncthreads();
    }
    if (P.debug && blockIdx.x == 0 && tid == 0)
        printf("[impute cycles] bwd-vector %lld  bwd-gram %lld  fwd-known %lld  fwd-Ad %lld  rho %lld  pdf %lld  select %lld\n", tm[0], tm[1],
               tm[2], tm[3], tm[4], tm[5], tm[6]);
#undef TICK
#undef TOCK
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
                 const double* xgrid, int G, const double* uniforms, int64_t uniforms_per_instance, int n_traj,
                 const mpst_impute_opts* io, double* out, double* err_out) {
    const double max_jump = io ? io->max_jump : -1.0;
    const bool backwards = io && io->backwards;
    const bool rejection = io && method == MPST_IMPUTE_ITS && io->rejection_threshold >= 0.0;
    if (rejection && io->max_trials < 1) { c->err = "impute_batch: max_trials must be positive"; return MPST_E_INVALID; }
    if (!X || !missing || !xgrid || !out || n < 0 || G < 2 || c->T == 0) { c->err = "impute_batch: bad arguments"; return MPST_E_INVALID; }
    if (class_idx < 0 || class_idx >= c->C) { c->err = "impute_batch: class index out of range"; return MPST_E_INVALID; }
    if (method < MPST_IMPUTE_MEDIAN || method > MPST_IMPUTE_ITS) { c->err = "impute_batch: unknown method"; return MPST_E_INVALID; }
    if (method == MPST_IMPUTE_ITS && !uniforms) { c->err = "impute_batch: ITS needs the uniform draws"; return MPST_E_INVALID; }
    // table encodings (projected Legendre, SL / SLTD, split bases): per-site grid states, values encoded with their site's table
    const bool tab = c->basis >= MPST_BASIS_TABLE_LEGENDRE_PROJ && c->basis <= MPST_BASIS_TABLE_SPLIT;
    if (tab && (c->enc.kind != c->basis || c->enc.d != c->d || !c->enc.ip || !c->enc.dp || (c->enc.nsites != 1 && c->enc.nsites != c->T))) {
        c->err = "impute_batch: no coefficient table set for this encoding (mpst_set_encoding_table)";
        return MPST_E_INVALID;
    }
    if (c->have_phi || (!tab && c->basis != MPST_BASIS_LEGENDRE_NO_NORM && c->basis != MPST_BASIS_LEGENDRE_NORM && c->basis != MPST_BASIS_UNIFORM)) {
        c->err = "impute_batch: needs one of the on-device real bases";
        return MPST_E_UNSUPPORTED;
    }
    if (n == 0) return MPST_OK;
    if (method != MPST_IMPUTE_ITS) n_traj = 1;
    if (n_traj < 1) n_traj = 1;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int T = c->T, d = c->d;
    // impute_order = :backwards (MPS_methods.jl:113-118: orthogonalize to the last site, walk right to left) is the
    // forward walk on the mirrored chain: position p holds site T-1-p with its two links exchanged; series, masks and
    // results are mirrored on the host.  pos(j) = chain position of site j.
    auto pos = [&](int j) { return backwards ? T - 1 - j : j; };
    std::vector<int> chi(T + 1, 1);
    std::vector<int64_t> off(T, 0);
    int chimax = 1;
    int64_t tot = 0;
    for (int p = 0; p < T; p++) {
        const Core& k = c->cores[pos(p)];
        if (!k.dev) { c->err = "impute_batch: cores not set"; return MPST_E_INVALID; }
        chi[p] = backwards ? k.chi_r : k.chi_l;
        chi[p + 1] = backwards ? k.chi_l : k.chi_r;
        chimax = std::max(chimax, std::max(k.chi_l, k.chi_r));
        off[p] = tot;
        tot += (int64_t)d * k.chi_l * k.chi_r;
    }
    std::vector<double> Xrev;
    std::vector<uint8_t> Mrev;
    if (backwards) {
        Xrev.resize((size_t)n * T);
        Mrev.resize((size_t)n * T);
        for (int64_t i = 0; i < n; i++)
            for (int j = 0; j < T; j++) { Xrev[i * T + j] = X[i * T + T - 1 - j]; Mrev[i * T + j] = missing[i * T + T - 1 - j]; }
        X = Xrev.data();
        missing = Mrev.data();
    }
    const auto t_begin = std::chrono::steady_clock::now();
    const int n8 = (chimax + 7) & ~7, ld = n8 + 4;
    size_t smem = sizeof(double) * ((size_t)4 * n8 * ld + 2 * n8 + 2 * (size_t)d * ld + 2 * (size_t)d * d + MPST_MAX_D + 32 +
                                    std::max(3 * NT + 16, 4 * n8) + MPST_MAX_D * MPST_MAX_D);
    const bool dbuf = smem + sizeof(double) * (size_t)n8 * ld <= 227 * 1024 && !c->flag[F_IMPUTE_NODBUF];
    if (dbuf) smem += sizeof(double) * (size_t)n8 * ld;
    if (smem > 227 * 1024) { c->err = "impute_batch: chi too large for the shared-memory Gram matrices (chi <= 72 at d = 16)"; return MPST_E_UNSUPPORTED; }
    // host-side Kmax
    int Kmax = 0;
    for (int64_t i = 0; i < n; i++) { int k = 0; for (int j = 0; j < T; j++) k += missing[i * T + j] ? 1 : 0; Kmax = std::max(Kmax, k); }
    if (err_out) memset(err_out, 0, sizeof(double) * (size_t)n * n_traj * T);
    if (Kmax == 0) {
        for (int64_t i = 0; i < n; i++)
            for (int tr = 0; tr < n_traj; tr++)
                for (int j = 0; j < T; j++) out[(i * n_traj + tr) * T + j] = X[i * T + pos(j)];
        return MPST_OK;
    }
    const int64_t ustride = rejection ? uniforms_per_instance : (int64_t)n_traj * Kmax;
    if (rejection && ustride < (int64_t)n_traj * Kmax * io->max_trials) {
        c->err = "impute_batch: rejection sampling needs n_traj * K_max * max_trials uniforms per instance";
        return MPST_E_INVALID;
    }
    const int grid = (int)std::min<int64_t>(n, (int64_t)c->sm_count);       // one resident CTA per SM (shared memory), instances strided
    double *dcores = nullptr, *dX = nullptr, *dgrid = nullptr, *dgenc = nullptr, *dunif = nullptr, *dout = nullptr, *dGR = nullptr, *dp = nullptr;
    double* derr = nullptr;
    uint8_t* dmask = nullptr;
    int64_t* doff = nullptr;
    int* dchi = nullptr;
    int rc = MPST_OK;
    auto cleanup = [&]() {};                                       // buffers belong to the context (freed by mpst_destroy)
#define IMP_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { c->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cleanup(); return MPST_E_CUDA; } } while (0)
    // grow-only work buffers held by the context: a second call of the same shape pays for copies only
    auto reserve = [&](int slot, size_t bytes, void** out) -> cudaError_t {
        if (c->imp_ptr[slot] && c->imp_cap[slot] >= bytes) { *out = c->imp_ptr[slot]; return cudaSuccess; }
        if (c->imp_ptr[slot]) { cudaFree(c->imp_ptr[slot]); c->imp_ptr[slot] = nullptr; c->imp_cap[slot] = 0; }
        cudaError_t e = cudaMalloc(&c->imp_ptr[slot], bytes);
        if (e == cudaSuccess) { c->imp_cap[slot] = bytes; *out = c->imp_ptr[slot]; }
        return e;
    };
    IMP_TRY(reserve(0, sizeof(double) * tot, (void**)&dcores));
    IMP_TRY(reserve(1, sizeof(double) * n * T, (void**)&dX));
    IMP_TRY(reserve(2, (size_t)n * T, (void**)&dmask));
    IMP_TRY(reserve(3, sizeof(double) * G, (void**)&dgrid));
    const bool per_site = tab && c->enc.nsites != 1;                // one block of grid states per chain position
    IMP_TRY(reserve(4, sizeof(double) * (size_t)G * d * (per_site ? T : 1), (void**)&dgenc));
    IMP_TRY(reserve(5, sizeof(double) * n * n_traj * T, (void**)&dout));
    IMP_TRY(reserve(6, sizeof(double) * ((size_t)grid * Kmax * chimax * chimax + 64), (void**)&dGR));
    IMP_TRY(reserve(7, sizeof(double) * (size_t)grid * G, (void**)&dp));
    IMP_TRY(reserve(8, sizeof(int64_t) * T, (void**)&doff));
    IMP_TRY(reserve(9, sizeof(int) * (T + 1), (void**)&dchi));
    if (uniforms) {
        IMP_TRY(reserve(10, sizeof(double) * n * ustride, (void**)&dunif));
        IMP_TRY(cudaMemcpyAsync(dunif, uniforms, sizeof(double) * n * ustride, cudaMemcpyHostToDevice, c->stream));
    }
    if (err_out) {
        IMP_TRY(reserve(11, sizeof(double) * n * n_traj * T, (void**)&derr));
        IMP_TRY(cudaMemsetAsync(derr, 0, sizeof(double) * n * n_traj * T, c->stream));
    }
    IMP_TRY(cudaMemcpyAsync(dX, X, sizeof(double) * n * T, cudaMemcpyHostToDevice, c->stream));
    IMP_TRY(cudaMemcpyAsync(dmask, missing, (size_t)n * T, cudaMemcpyHostToDevice, c->stream));
    IMP_TRY(cudaMemcpyAsync(dgrid, xgrid, sizeof(double) * G, cudaMemcpyHostToDevice, c->stream));
    IMP_TRY(cudaMemcpyAsync(doff, off.data(), sizeof(int64_t) * T, cudaMemcpyHostToDevice, c->stream));
    IMP_TRY(cudaMemcpyAsync(dchi, chi.data(), sizeof(int) * (T + 1), cudaMemcpyHostToDevice, c->stream));
    for (int p = 0; p < T; p++) {
        const Core& k = c->cores[pos(p)];
        CoreView v;
        v.p = k.dev; v.ss = 1;
        if (k.orient == ORIENT_LEFT) { v.sa = d; v.sb = (long)d * k.chi_l; } else { v.sb = d; v.sa = (long)d * k.chi_r; }
        v.sc = k.has_label ? (long)d * k.chi_l * k.chi_r : 0;
        if (backwards) std::swap(v.sa, v.sb);                      // mirrored chain: the links change sides
        const int64_t ne = (int64_t)d * k.chi_l * k.chi_r;
        slice_core_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, c->stream>>>(v, d, chi[p], chi[p + 1], k.has_label ? class_idx : 0, dcores + off[p]);
        c->launches++;
    }
    // grid states, imputation.jl:92-107 (time-dependent encodings: one set per site, in chain order)
    if (per_site) {
        for (int p = 0; p < T && rc == MPST_OK; p++) rc = launch_encode_site(c, pos(p), dgrid, G, dgenc + (size_t)p * G * d, d);
    } else rc = launch_encode_site(c, 0, dgrid, G, dgenc, d);
    if (rc != MPST_OK) { cleanup(); return rc; }
    ImpParams P;
    P.cores = dcores; P.core_off = doff; P.chi = dchi; P.X = dX; P.mask = dmask; P.grid = dgrid; P.genc = dgenc;
    P.uniforms = dunif; P.out = dout; P.gr_scratch = dGR; P.p_scratch = dp;
    P.T = T; P.d = d; P.G = G; P.ntraj = n_traj; P.Kmax = Kmax; P.chimax = chimax; P.method = method; P.basis = c->basis;
    P.n = n; P.max_jump = max_jump;
    P.err = derr; P.get_err = (io && io->get_err) ? 1 : 0; P.max_trials = io ? io->max_trials : 0;
    P.rej_thr = rejection ? io->rejection_threshold : -1.0; P.ustride = ustride;
    P.debug = c->flag[F_IMPUTE_DEBUG] ? 1 : 0;
    P.genc_stride = per_site ? (int64_t)G * d : 0;
    P.tab_kind = tab ? c->enc.kind : 0; P.tab_nsites = tab ? c->enc.nsites : 1; P.mirror = backwards ? 1 : 0;
    P.tab_istride = c->enc.istride; P.tab_dstride = c->enc.dstride; P.tab_ip = c->enc.ip; P.tab_dp = c->enc.dp;
    P.dbuf = dbuf ? 1 : 0;
    {
        double ha[MPST_MAX_D], hb[MPST_MAX_D];
        ha[0] = hb[0] = 0.0;
        for (int l = 1; l < MPST_MAX_D; l++) { ha[l] = (double)(2 * l - 1) / (double)l; hb[l] = (double)(l - 1) / (double)l; }
        IMP_TRY(cudaMemcpyToSymbolAsync(c_leg_a, ha, sizeof(ha), 0, cudaMemcpyHostToDevice, c->stream));
        IMP_TRY(cudaMemcpyToSymbolAsync(c_leg_b, hb, sizeof(hb), 0, cudaMemcpyHostToDevice, c->stream));
        IMP_TRY(cudaStreamSynchronize(c->stream));                                 // ha/hb are stack temporaries
    }
    P.leg_pq = nullptr; P.leg_wq = nullptr; P.nq = 0;
    if ((c->basis == MPST_BASIS_LEGENDRE_NO_NORM || c->basis == MPST_BASIS_LEGENDRE_NORM) && d >= 8 && !c->flag[F_IMPUTE_NOSERIES]) {
        // Chebyshev-Gauss nodes x_i = cos(pi (i + 1/2) / nq), nq = 2d-1; P_l at the nodes and the analysis matrix
        // wq[k][i] = (k == 0 ? 1 : 2) / nq * T_k(x_i) (discrete orthogonality: exact for polynomials of degree < nq)
        const int nq = 2 * d - 1;
        std::vector<double> xs(nq), tab((size_t)d * nq + (size_t)nq * nq);
        const double pi = 3.14159265358979323846;
        for (int i = 0; i < nq; i++) xs[i] = cos(pi * (i + 0.5) / nq);
        double* pq = tab.data();
        double* wq = tab.data() + (size_t)d * nq;
        for (int i = 0; i < nq; i++) {
            double p0 = 1.0, p1 = xs[i];
            for (int l = 0; l < d; l++) {
                const double pl = l == 0 ? 1.0 : l == 1 ? xs[i] : ((2 * l - 1) * xs[i] * p1 - (l - 1) * p0) / l;
                if (l >= 2) { p0 = p1; p1 = pl; }
                pq[(size_t)l * nq + i] = pl;
            }
            for (int k = 0; k < nq; k++) wq[(size_t)k * nq + i] = (k == 0 ? 1.0 : 2.0) / nq * cos(pi * k * (i + 0.5) / nq);
        }
        double* dtab = nullptr;
        IMP_TRY(reserve(12, sizeof(double) * tab.size(), (void**)&dtab));
        IMP_TRY(cudaMemcpyAsync(dtab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, c->stream));
        IMP_TRY(cudaStreamSynchronize(c->stream));                 // tab is a local temporary
        P.leg_pq = dtab; P.leg_wq = dtab + (size_t)d * nq; P.nq = nq;
    }
    IMP_TRY(cudaFuncSetAttribute(impute_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IMP_TRY(cudaFuncSetAttribute(impute_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const auto t_setup = std::chrono::steady_clock::now();
    prof_begin(c, MPST_T_IMPUTE);
    if (tab) impute_kernel<true><<<grid, NT, smem, c->stream>>>(P);
    else impute_kernel<false><<<grid, NT, smem, c->stream>>>(P);
    prof_end(c, MPST_T_IMPUTE);
    c->launches++;
    IMP_TRY(cudaGetLastError());
    IMP_TRY(cudaMemcpyAsync(out, dout, sizeof(double) * n * n_traj * T, cudaMemcpyDeviceToHost, c->stream));
    if (err_out) IMP_TRY(cudaMemcpyAsync(err_out, derr, sizeof(double) * n * n_traj * T, cudaMemcpyDeviceToHost, c->stream));
    IMP_TRY(cudaStreamSynchronize(c->stream));
    if (backwards) {
        for (int64_t r = 0; r < n * n_traj; r++) {
            std::reverse(out + r * T, out + (r + 1) * T);
            if (err_out) std::reverse(err_out + r * T, err_out + (r + 1) * T);
        }
    }
    if (P.debug) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "[impute host] setup %.1f ms, kernel + copy back %.1f ms (grid %d, n %lld)\n",
                1e3 * std::chrono::duration<double>(t_setup - t_begin).count(),
                1e3 * std::chrono::duration<double>(t_end - t_setup).count(), grid, (long long)n);
    }
#undef IMP_TRY
    cleanup();
    return MPST_OK;
}
