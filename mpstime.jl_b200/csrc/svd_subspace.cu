// K5, fast path: truncated SVD of the bond matrix by block subspace iteration + Rayleigh-Ritz.
//
// Only the chi_max largest singular triplets of M (m x n) survive decomposeBT's truncation
// (reference Training/RealRealHighDimension.jl:146-203) and the NDTensors rule needs nothing of the rest but
// its total weight ||M||_F^2 - sum(kept sigma^2).  The singular values of a trained bond tensor decay
// geometrically (measured: sigma_{2k}/sigma_k ~ 1e-1..1e-2 at k = 40), so a subspace of p = 2k vectors
// converges at (sigma_{p+1}/sigma_k)^2 per iteration: 5 iterations reach rounding level where the full
// one-sided Jacobi needs 20-30 latency-bound sweeps.
//   Q <- orth(random n x p)
//   repeat:  Z <- orth(M Q);  Q <- orth(M^T Z)              (orth = Cholesky-QR, twice on the last pass)
//   H = (M Q)^T (M Q)  ->  W Lambda W^T   (two-sided Jacobi on the p x p matrix in one CTA)
//   sigma_i = sqrt(Lambda_i),  V_k = Q W_k  (orthonormal core),  U_k Sigma_k = (M Q) W_k  (moving core)
// Every step is a DMMA GEMM or a single-CTA kernel on a p x p matrix.  An a-posteriori residual
// ||M^T (M v_i) - sigma_i^2 v_i|| / sigma_1^2 is checked on the device; if it is not at rounding level
// (or a Cholesky pivot breaks down: numerically rank-deficient block) the caller falls back to the exact
// full Jacobi SVD (svd_jacobi.cu), so the result never depends on the spectrum being friendly.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "mpst_common.cuh"

int launch_dgemm(mpst_ctx* c, int ta, int tb, int M, int N, int K, const double* A, int64_t lda, const double* B,
                 int64_t ldb, double* C, int64_t ldc);
int launch_dgemm_splitk(mpst_ctx* c, int ta, int tb, int M, int N, int K, const double* A, int64_t lda, const double* B,
                        int64_t ldb, double* Cparts, int max_splits, int* splits_out);

namespace {

__device__ __forceinline__ double hash_unit(uint64_t x) {          // splitmix64 -> (-1, 1)
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

__global__ void rand_init_kernel(double* __restrict__ Q, int n, int p, int64_t ld, uint64_t seed) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * p) return;
    const int j = (int)(e / n), i = (int)(e - (int64_t)j * n);
    Q[i + ld * j] = hash_unit(seed + (uint64_t)e * 0x632BE59BD9B4E019ull);
}

// G = sum of `splits` partial Gram matrices (p x p, symmetric positive definite) = L L^T;
// Rinv = (L^T)^-1 = (L^-1)^T (upper, column-major p x p).  Single CTA, one right-looking pass that factors
// column k and immediately eliminates it from the running inverse (forward substitution on the identity),
// two block barriers per column, every update a 2-D thread-parallel rank-1 update in shared memory.
// status[0] |= 1 when a pivot is not safely positive (caller falls back to the exact SVD).
__global__ void __launch_bounds__(256)
chol_inv_kernel(const double* __restrict__ G, int splits, int p, double* __restrict__ Rinv, int* __restrict__ status) {
    extern __shared__ double sm[];
    const int ld = p + 1;
    double* A = sm;                                    // [p][ld]: lower triangle -> L
    double* X = sm + (size_t)p * ld;                   // [p][ld]: running B (rows > k) / finished X = L^-1 (rows <= k)
    __shared__ int s_bad;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // sum the split-K partials: partial-major so each thread has p*p/256 independent loads in flight
    for (int e = tid; e < p * p; e += 256) {
        const int i = e % p, j = e / p;
        A[i * ld + j] = G[e];
        X[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    for (int z = 1; z < splits; z++)
        for (int e = tid; e < p * p; e += 256) A[(e % p) * ld + e / p] += G[(size_t)z * p * p + e];
    if (tid == 0) s_bad = 0;
    __syncthreads();
    double tr = 0.0;
    for (int i = 0; i < p; i++) tr = fmax(tr, A[i * ld + i]);
    const double tiny = 1e-13 * tr;                    // kappa(G) beyond ~1e13: CholeskyQR no longer trustworthy
    for (int k = 0; k < p; k++) {
        const double dk = A[k * ld + k];
        if (tid == 0 && !(dk > tiny)) s_bad = 1;
        const double inv = rsqrt(fmax(dk, tiny));      // 1 / l_kk
        __syncthreads();                               // everybody has read the pivot
        // scale column k of L (rows >= k) and finish row k of X (columns <= k)
        for (int i = k + tid; i < p; i += 256) A[i * ld + k] *= inv;      // A[k][k] = dk/sqrt(dk) = l_kk
        for (int j = tid; j <= k; j += 256) X[k * ld + j] *= inv;
        __syncthreads();
        // rank-1 updates with column k:  A[i][l] -= L[i][k] L[l][k] (k < l <= i),  B[i][j] -= L[i][k] X[k][j] (j <= k < i)
        // operands are gathered into registers first so the shared-memory latencies overlap (p <= 112 -> <= 7 per loop)
        for (int i = k + 1 + ty; i < p; i += 16) {
            const double lik = A[i * ld + k];
            double av[7], lv[7], xv[7], kv[7];
#pragma unroll
            for (int t = 0; t < 7; t++) {
                const int l = k + 1 + tx + 16 * t;
                if (l <= i) { av[t] = A[i * ld + l]; lv[t] = A[l * ld + k]; }
                const int j = tx + 16 * t;
                if (j <= k) { xv[t] = X[i * ld + j]; kv[t] = X[k * ld + j]; }
            }
#pragma unroll
            for (int t = 0; t < 7; t++) {
                const int l = k + 1 + tx + 16 * t;
                if (l <= i) A[i * ld + l] = av[t] - lik * lv[t];
                const int j = tx + 16 * t;
                if (j <= k) X[i * ld + j] = xv[t] - lik * kv[t];
            }
        }
        // no barrier needed before the next pivot read: A[k+1][k+1] is written by its owner above and the
        // barrier at the top of the next iteration orders it (pivot is read before that barrier -> add one)
        __syncthreads();
    }
    for (int e = tid; e < p * p; e += 256) {
        const int jj = e % p, i = e / p;               // Rinv[jj + p*i] = X[i][jj]  (i >= jj)
        Rinv[e] = (i >= jj) ? X[i * ld + jj] : 0.0;
    }
    if (tid == 0 && s_bad) atomicOr(status, 1);
}

// Eigen-decomposition of a symmetric q x q matrix (embedded in an even order p >= q) by cyclic two-sided
// Jacobi in shared memory, one CTA.  Per round the p/2 disjoint rotations are computed first, then every
// 2x2 block (pair I, pair J), I <= J, of A is transformed once with both rotations (B' = R_I^T B R_J) and
// mirrored, and V's column pairs are rotated: one pass over A (upper half) and V per round, two barriers.
// W: column-major eigenvectors, ev: eigenvalues (unsorted).  status[1] = sweeps used.
__global__ void __launch_bounds__(1024)
sym_eig_kernel(const double* __restrict__ H, int splits, int q, int p, double* __restrict__ W, double* __restrict__ ev,
               int* __restrict__ status) {
    extern __shared__ double sm[];
    const int ld = p + 1, hp = p / 2;
    double* A = sm;
    double* V = sm + (size_t)p * ld;
    double* rc = V + (size_t)p * ld;                   // [hp] cos
    double* rs = rc + hp;                              // [hp] sin
    int* rp = reinterpret_cast<int*>(rs + hp);         // [hp] p index, [hp] q index
    int* rq = rp + hp;
    __shared__ int any_rot;
    __shared__ unsigned long long s_maxaa;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int e = tid; e < p * p; e += nthr) {
        const int i = e / p, j = e % p;
        A[i * ld + j] = 0.0;
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int z = 0; z < splits; z++)                      // partial-major: independent loads in flight; each (i,j) has
        for (int e = tid; e < q * q; e += nthr) {         // one owner thread and a fixed summation order (deterministic)
            const int i = e % q, j = e / q;
            A[i * ld + j] += 0.5 * (H[(size_t)z * q * q + e] + H[(size_t)z * q * q + j + (size_t)q * i]);
        }
    __syncthreads();
    double tr = 0.0;
    for (int i = 0; i < p; i++) tr += A[i * ld + i];
    // H is a Gram matrix: its entries carry absolute errors ~eps*trace, so couplings below a few eps*trace are
    // noise (rotating them never terminates); the eigenvalues are resolved to that absolute accuracy anyway.
    const double floor_abs = 1e-15 * tr;
    int sweep = 0;
    for (; sweep < 60; sweep++) {
        if (tid == 0) { any_rot = 0; s_maxaa = 0ull; }
        __syncthreads();
        for (int rd = 0; rd < p - 1; rd++) {
            if (tid < hp) {
                const int k = tid;
                int a, b;
                if (k == 0) { a = p - 1; b = rd; }
                else { a = (rd + k) % (p - 1); b = (rd - k + (p - 1)) % (p - 1); }
                const int pp = min(a, b), qq = max(a, b);
                const double app = A[pp * ld + pp], aqq = A[qq * ld + qq], apq = A[pp * ld + qq];
                const double aa = fabs(apq);
                double c = 1.0, sn = 0.0;
                if (aa > floor_abs && aa * aa > 1e-30 * fabs(app * aqq)) {
                    // cos/sin of the double angle, then half-angle: two rsqrt, no division
                    const double zeta = aqq - app, beta = 2.0 * apq;
                    const double inv_r = rsqrt(zeta * zeta + beta * beta);
                    const double cos2 = fabs(zeta) * inv_r, sin2 = (zeta >= 0.0 ? beta : -beta) * inv_r;
                    const double c2 = 0.5 + 0.5 * cos2;
                    const double inv_c = rsqrt(c2);
                    c = c2 * inv_c;
                    sn = 0.5 * sin2 * inv_c;
                    any_rot = 1;
                    atomicMax(&s_maxaa, (unsigned long long)__double_as_longlong(aa));
                }
                rc[k] = c; rs[k] = sn; rp[k] = pp; rq[k] = qq;
            }
            __syncthreads();
            // A blocks (I <= J), mirrored.  The upper triangle is folded into a (hp/2) x (hp+1) rectangle (rows I and
            // hp-1-I share a line) so every thread gets a block; odd hp walks the square and skips I > J.
            const int nblk = (hp & 1) ? hp * hp : (hp / 2) * (hp + 1);
            for (int e = tid; e < nblk; e += nthr) {
                int I, J;
                if (hp & 1) { I = e / hp; J = e - I * hp; if (I > J) continue; }
                else {
                    const int I0 = e / (hp + 1), t = e - I0 * (hp + 1);
                    if (t < hp - I0) { I = I0; J = I0 + t; }
                    else { I = hp - 1 - I0; J = I + (t - (hp - I0)); }
                }
                const int pi = rp[I], qi = rq[I], pj = rp[J], qj = rq[J];
                const double ci = rc[I], si = rs[I], cj = rc[J], sj = rs[J];
                const double b11 = A[pi * ld + pj], b12 = A[pi * ld + qj], b21 = A[qi * ld + pj], b22 = A[qi * ld + qj];
                // columns: B R_J
                const double t11 = cj * b11 - sj * b12, t12 = sj * b11 + cj * b12;
                const double t21 = cj * b21 - sj * b22, t22 = sj * b21 + cj * b22;
                // rows: R_I^T (.)
                const double n11 = ci * t11 - si * t21, n12 = ci * t12 - si * t22;
                const double n21 = si * t11 + ci * t21, n22 = si * t12 + ci * t22;
                A[pi * ld + pj] = n11; A[pi * ld + qj] = n12; A[qi * ld + pj] = n21; A[qi * ld + qj] = n22;
                if (I != J) { A[pj * ld + pi] = n11; A[qj * ld + pi] = n12; A[pj * ld + qi] = n21; A[qj * ld + qi] = n22; }
            }
            // V column pairs (4 independent element pairs in flight per thread)
            for (int e0 = tid; e0 < hp * p; e0 += 4 * nthr) {
                double vx[4], vy[4], cc[4], ss[4];
                int ip[4], iq[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int e = e0 + t * nthr;
                    if (e < hp * p) {
                        const int k = e / p, r = e - k * p;
                        ip[t] = r * ld + rp[k]; iq[t] = r * ld + rq[k];
                        cc[t] = rc[k]; ss[t] = rs[k];
                        vx[t] = V[ip[t]]; vy[t] = V[iq[t]];
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int e = e0 + t * nthr;
                    if (e < hp * p) {
                        V[ip[t]] = cc[t] * vx[t] - ss[t] * vy[t];
                        V[iq[t]] = ss[t] * vx[t] + cc[t] * vy[t];
                    }
                }
            }
            __syncthreads();
        }
        // quadratic convergence: a sweep whose largest rotated entry was <= 1e-9*trace leaves <= ~1e-18*trace
        if (!any_rot || __longlong_as_double((long long)s_maxaa) <= 1e-9 * tr) break;
        __syncthreads();
    }
    if (tid == 0 && sweep >= 60) atomicOr(status, 2);
    if (tid == 0) status[1] = sweep;
    for (int e = tid; e < p * p; e += nthr) {
        const int i = e / p, j = e % p;
        W[i + (size_t)p * j] = V[i * ld + j];
    }
    for (int i = tid; i < p; i += nthr) ev[i] = A[i * ld + i];
}

// rank the p Ritz values, apply the NDTensors truncation rule with the weight outside the subspace
// (trace - sum) already discarded; perm[k] = column of the k-th largest, Psorted, iscal[0] = chi_new.
__global__ void __launch_bounds__(256)
ritz_trunc_kernel(const double* __restrict__ ev, int p, int n_total, const double* __restrict__ trace_dev, int maxdim,
                  double cutoff, int* __restrict__ perm, double* __restrict__ Psorted, int* __restrict__ iscal) {
    for (int j = threadIdx.x; j < p; j += blockDim.x) {
        const double pj = ev[j];
        int rank = 0;
        for (int q = 0; q < p; q++) {
            const double pq = ev[q];
            rank += (pq > pj) || (pq == pj && q < j);
        }
        perm[rank] = j;
        Psorted[rank] = fmax(pj, 0.0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
        for (int q = 0; q < p; q++) sum += Psorted[q];
        double scale = trace_dev ? *trace_dev : 1.0;
        if (!(scale > 0.0)) scale = 1.0;
        double err = fmax(scale - sum, 0.0);               // everything outside the Ritz subspace
        (void)n_total;
        int keep = p;
        while (keep > maxdim) { err += Psorted[keep - 1]; keep--; }
        while (keep > 1 && err + Psorted[keep - 1] <= cutoff * scale) { err += Psorted[keep - 1]; keep--; }
        iscal[0] = max(keep, 1);
    }
}

// Wk[:, kk] = W[:, perm[kk]]   (p x kmax, column-major)
__global__ void gather_cols_kernel(const double* __restrict__ W, int p, const int* __restrict__ perm, int kmax,
                                   double* __restrict__ Wk) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p * kmax) return;
    const int kk = e / p, i = e % p;
    Wk[e] = W[i + (size_t)p * perm[kk]];
}

// label core [c][x + Dx*kk] <- UkSk (m x k column-major, rows r = c*Dx + x)
__global__ void scatter_label_kernel(const double* __restrict__ UkSk, int m, int Dx, const int* __restrict__ iscal,
                                     double* __restrict__ label_core) {
    const int chi = iscal[0];
    const int64_t tot = (int64_t)m * chi;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e / m), r = (int)(e % m);
        const int cls = r / Dx, xx = r - cls * Dx;
        label_core[(size_t)cls * Dx * chi + xx + (size_t)Dx * kk] = UkSk[e];
    }
}

// res = max_i || T2[:, i] - P_i * Vk[:, i] || / P_0   over the kept columns
__global__ void __launch_bounds__(256)
residual_kernel(const double* __restrict__ T2, const double* __restrict__ Vk, const double* __restrict__ Psorted,
                const int* __restrict__ iscal, int n, unsigned long long* __restrict__ out_bits) {
    __shared__ double sh[8];
    const int i = blockIdx.x;
    if (i >= iscal[0]) return;
    const double lam = Psorted[i];
    double s = 0.0;
    for (int r = threadIdx.x; r < n; r += 256) {
        const double d = T2[r + (size_t)n * i] - lam * Vk[r + (size_t)n * i];
        s += d * d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        const double rel = sqrt(t) / fmax(Psorted[0], 1e-300);
        atomicMax(out_bits, (unsigned long long)__double_as_longlong(rel));
    }
}
// scale the columns of Vk (n x k): Vk[:, i] *= 1/sqrt(P_i)   (wide Gram path: v_i = M^T u_i / sigma_i)
__global__ void scale_cols_kernel(double* __restrict__ Vk, int n, const double* __restrict__ Psorted,
                                  const int* __restrict__ iscal) {
    const int chi = iscal[0];
    const int64_t tot = (int64_t)n * chi;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / n);
        Vk[e] *= rsqrt(fmax(Psorted[i], 1e-300));
    }
}
// Uk[:, i] *= sqrt(P_i)   (wide Gram path: moving core = U_k Sigma_k)
__global__ void scale_cols_sqrt_kernel(double* __restrict__ Uk, int m, const double* __restrict__ Psorted,
                                       const int* __restrict__ iscal) {
    const int chi = iscal[0];
    const int64_t tot = (int64_t)m * chi;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x)
        Uk[e] *= sqrt(fmax(Psorted[(int)(e / m)], 0.0));
}
}  // namespace

// M: column-major m x n (leading dimension ldm) on the device, already scaled.  trace_dev: ||M||_F^2 on the
// device or nullptr (== 1).  On success *done = true and the cores are written; *done = false means
// "not applicable / not converged": the caller must run the full Jacobi SVD.
// Three modes, all ending in the single-CTA symmetric eigen-solver on a matrix of order <= 112:
//   tall-Gram  (n <= 112)          : H = M^T M,  V = eigenvectors, U S = M V
//   wide-Gram  (m <= 112, m < n)   : H = M M^T,  U = eigenvectors, V = M^T U S^-1   (kept sigma >= 1e-5 sigma_1)
//   subspace   (otherwise)         : block subspace iteration, H = (M Q)^T (M Q)
// Working with sigma^2 resolves weights down to eps*sigma_1^2, so these paths are used only when the
// truncation cutoff is >= 1e-12 (the reference default is 1e-10); smaller cutoffs take the exact Jacobi.
int svd_subspace_device(mpst_ctx* c, const double* M, int64_t ldm, int m, int n, int C, int chi_max, double cutoff,
                        const double* trace_dev, double* label_core, double* ortho_core, int* chi_new,
                        double* sigma_host, bool* done) {
    *done = false;
    if (getenv("MPST_SVD_NOSUB") || cutoff < 1e-12) return MPST_OK;
    const int k = std::min(chi_max, std::min(m, n));
    const int PMAX = 112, MAXSPLIT = 16;
    int mode;                                                          // 0 tall-Gram, 1 wide-Gram, 2 subspace
    int p;
    if (n <= PMAX) { mode = 0; p = n + (n & 1); }
    else if (m <= PMAX && m < n) { mode = 1; p = m + (m & 1); }
    else {
        mode = 2;
        // measured on trained bonds (k = 40): p = 80 with 5 iterations beats p = 96 with 4 and p = 112 with 3 --
        // the p^3 single-CTA kernels (Cholesky, Rayleigh-Ritz eigen-solver) dominate, not the GEMMs
        p = std::min((int)round_up(2 * k + (getenv("MPST_SVD_OVS") ? atoi(getenv("MPST_SVD_OVS")) : 0), 16), PMAX);
        if (p < k + 32 || n <= p || m < p) return MPST_OK;            // too little oversampling: full Jacobi
    }
    const size_t need = (size_t)2 * n * p + (size_t)2 * m * p + (size_t)(MAXSPLIT + 3) * p * p + (size_t)n * k +
                        (size_t)m * k + 4 * p + 64;
    TRY(ensure_buf(c, &c->sub, &c->subcap, need));
    double* Qa = c->sub;
    double* Qb = Qa + (size_t)n * p;
    double* Za = Qb + (size_t)n * p;
    double* Zb = Za + (size_t)m * p;
    double* Gm = Zb + (size_t)m * p;                                   // MAXSPLIT partial Gram matrices
    double* Ri = Gm + (size_t)MAXSPLIT * p * p;
    double* Wm = Ri + (size_t)p * p;
    double* Wk = Wm + (size_t)p * p;
    double* T2 = Wk + (size_t)p * p;                                   // n x k
    double* Uk = T2 + (size_t)n * k;                                   // m x k
    double* ev = Uk + (size_t)m * k;                                   // p, then Psorted p
    int* status = c->iscal + 8;
    unsigned long long* resbits = reinterpret_cast<unsigned long long*>(c->scal + 10);
    const size_t chol_smem = 2 * sizeof(double) * (size_t)p * (p + 1);
    const size_t eig_smem = 2 * sizeof(double) * (size_t)p * (p + 1) + sizeof(double) * p + sizeof(int) * p + 16;
    CUDA_TRY(c, cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem));
    CUDA_TRY(c, cudaFuncSetAttribute(sym_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig_smem));
    CUDA_TRY(c, cudaMemsetAsync(status, 0, sizeof(int), c->stream));
    const bool dbg = getenv("MPST_SVD_DEBUG") != nullptr;
    if (dbg) cudaStreamSynchronize(c->stream);
    const auto t0 = std::chrono::steady_clock::now();
    const unsigned eig_threads = p >= 64 ? 1024u : 256u;

    auto finish = [&](const char* what, int iters, double res) -> int {
        *chi_new = c->hiscal[0];
        scatter_label_kernel<<<2 * c->sm_count, 256, 0, c->stream>>>(Uk, m, m / C, c->iscal, label_core);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
        if (sigma_host) {
            std::vector<double> tmp(*chi_new);
            CUDA_TRY(c, cudaMemcpy(tmp.data(), ev + p, sizeof(double) * (*chi_new), cudaMemcpyDeviceToHost));
            for (int q = 0; q < *chi_new; q++) sigma_host[q] = sqrt(tmp[q]);
        }
        if (dbg) {
            cudaStreamSynchronize(c->stream);
            fprintf(stderr, "[svd %s] m=%d n=%d p=%d iters=%d chi=%d residual=%.2e eigsweeps=%d t=%.3f ms\n", what, m, n, p, iters, *chi_new, res, c->hiscal[9],
                    1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        *done = true;
        return MPST_OK;
    };

    if (mode != 2) {
        // ---- Gram paths: one small symmetric eigenproblem, no iteration ----
        const int q = mode == 0 ? n : m;                               // order of H (p = q rounded up to even)
        int splits = 1;
        if (mode == 0) TRY(launch_dgemm_splitk(c, 1, 0, q, q, m, M, ldm, M, ldm, Gm, MAXSPLIT, &splits));   // M^T M
        else TRY(launch_dgemm_splitk(c, 0, 1, q, q, n, M, ldm, M, ldm, Gm, MAXSPLIT, &splits));            // M M^T
        sym_eig_kernel<<<1, eig_threads, eig_smem, c->stream>>>(Gm, splits, q, p, Wm, ev, status);
        ritz_trunc_kernel<<<1, 256, 0, c->stream>>>(ev, p, q, trace_dev, chi_max, cutoff, c->perm, ev + p, c->iscal);
        gather_cols_kernel<<<(p * k + 255) / 256, 256, 0, c->stream>>>(Wm, p, c->perm, k, Wk);
        c->launches += 3;
        if (mode == 0) {
            // V_k = W_k (rows < n), U_k S_k = M V_k
            CUDA_TRY(c, cudaMemcpy2DAsync(ortho_core, sizeof(double) * n, Wk, sizeof(double) * p, sizeof(double) * n, k,
                                          cudaMemcpyDeviceToDevice, c->stream));
            TRY(launch_dgemm(c, 0, 0, m, k, n, M, ldm, ortho_core, n, Uk, m));
        } else {
            // U_k = W_k (rows < m);  V_k = M^T U_k / sigma;  U_k S_k = U_k * sigma
            CUDA_TRY(c, cudaMemcpy2DAsync(Uk, sizeof(double) * m, Wk, sizeof(double) * p, sizeof(double) * m, k,
                                          cudaMemcpyDeviceToDevice, c->stream));
            TRY(launch_dgemm(c, 1, 0, n, k, m, M, ldm, Uk, m, ortho_core, n));
        }
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal, c->iscal, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal + 8, status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (c->hiscal[8] != 0) return MPST_OK;
        if (mode == 1) {
            scale_cols_kernel<<<2 * c->sm_count, 256, 0, c->stream>>>(ortho_core, n, ev + p, c->iscal);
            scale_cols_sqrt_kernel<<<2 * c->sm_count, 256, 0, c->stream>>>(Uk, m, ev + p, c->iscal);
            c->launches += 2;
        }
        return finish(mode == 0 ? "gram-tall" : "gram-wide", 0, 0.0);
    }

    // ---- subspace iteration ----
    // X (rows x p, ld = rows) -> orthonormal columns in `out`
    auto cholqr = [&](double* X, double* out, int rows) -> int {
        int splits = 1;
        TRY(launch_dgemm_splitk(c, 1, 0, p, p, rows, X, rows, X, rows, Gm, MAXSPLIT, &splits));
        chol_inv_kernel<<<1, 256, chol_smem, c->stream>>>(Gm, splits, p, Ri, status);
        c->launches++;
        TRY(launch_dgemm(c, 0, 0, rows, p, p, X, rows, Ri, p, out, rows));
        return MPST_OK;
    };
    {
        const int64_t tot = (int64_t)n * p;
        rand_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(Qa, n, p, n, 0x5DEECE66Dull + (uint64_t)m * 131 + n);
        c->launches++;
    }
    TRY(cholqr(Qa, Qb, n));
    TRY(cholqr(Qb, Qa, n));                                            // Q = Qa
    const int max_rounds = 3;
    int iters_done = 0;
    for (int round = 0; round < max_rounds; round++) {
        const int niter = round == 0 ? (getenv("MPST_SVD_IT") ? atoi(getenv("MPST_SVD_IT")) : (p >= 2 * k ? 5 : 7)) : 3;
        for (int it = 0; it < niter; it++) {
            TRY(launch_dgemm(c, 0, 0, m, p, n, M, ldm, Qa, n, Za, m));            // Z = M Q
            TRY(cholqr(Za, Zb, m));
            TRY(launch_dgemm(c, 1, 0, n, p, m, M, ldm, Zb, m, Qb, n));            // Y = M^T Z
            TRY(cholqr(Qb, Qa, n));
            iters_done++;
        }
        TRY(cholqr(Qa, Qb, n));                                                    // second pass: orthonormal to rounding
        CUDA_TRY(c, cudaMemcpyAsync(Qa, Qb, sizeof(double) * (size_t)n * p, cudaMemcpyDeviceToDevice, c->stream));
        // Rayleigh-Ritz
        int splits = 1;
        TRY(launch_dgemm(c, 0, 0, m, p, n, M, ldm, Qa, n, Za, m));                // Z = M Q
        TRY(launch_dgemm_splitk(c, 1, 0, p, p, m, Za, m, Za, m, Gm, MAXSPLIT, &splits));   // H = Z^T Z
        sym_eig_kernel<<<1, eig_threads, eig_smem, c->stream>>>(Gm, splits, p, p, Wm, ev, status);
        ritz_trunc_kernel<<<1, 256, 0, c->stream>>>(ev, p, n, trace_dev, chi_max, cutoff, c->perm, ev + p, c->iscal);
        gather_cols_kernel<<<(p * k + 255) / 256, 256, 0, c->stream>>>(Wm, p, c->perm, k, Wk);
        c->launches += 3;
        TRY(launch_dgemm(c, 0, 0, n, k, p, Qa, n, Wk, p, ortho_core, n));         // V_k = Q W_k
        TRY(launch_dgemm(c, 0, 0, m, k, p, Za, m, Wk, p, Uk, m));                 // U_k S_k = Z W_k
        // residual of the kept Ritz pairs: M^T (M v_i) - sigma_i^2 v_i
        TRY(launch_dgemm(c, 1, 0, n, k, m, M, ldm, Uk, m, T2, n));
        CUDA_TRY(c, cudaMemsetAsync(resbits, 0, sizeof(unsigned long long), c->stream));
        residual_kernel<<<k, 256, 0, c->stream>>>(T2, ortho_core, ev + p, c->iscal, n, resbits);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal, c->iscal, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal + 8, status, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->hscal + 10, resbits, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        const double res = c->hscal[10];
        if (c->hiscal[8] != 0 || !(res == res)) {
            if (dbg) fprintf(stderr, "[svd subspace] m=%d n=%d breakdown status=%d -> full Jacobi\n", m, n, c->hiscal[8]);
            return MPST_OK;
        }
        if (res <= 5e-14) return finish("subspace", iters_done, res);
        if (res > (round == 0 ? 1e-5 : 1e-10)) {                                   // spectrum too flat: full Jacobi
            if (dbg) fprintf(stderr, "[svd subspace] m=%d n=%d residual %.2e after %d its -> full Jacobi\n", m, n, res, iters_done);
            return MPST_OK;
        }
    }
    return MPST_OK;                                                                // not converged: full Jacobi
}
