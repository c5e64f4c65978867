// K5, fast path: truncated SVD of the bond matrix by block subspace iteration + Rayleigh-Ritz.
//
// Only the chi_max largest singular triplets of M (m x n) survive decomposeBT's truncation
// (reference Training/RealRealHighDimension.jl:146-203) and the NDTensors rule needs nothing of the rest but
// its total weight ||M||_F^2 - sum(kept sigma^2).  The singular values of a trained bond tensor decay
// geometrically (measured: sigma_{2k+16}/sigma_k ~ 1e-2 at k = 40), so a subspace of p = 2k+16 vectors
// converges at (sigma_{p+1}/sigma_k)^2 per iteration: 3-4 iterations reach rounding level where the full
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
#include <cstdio>
#include <cstdlib>
#include "mpst_common.cuh"

int launch_dgemm(mpst_ctx* c, int ta, int tb, int M, int N, int K, const double* A, int64_t lda, const double* B,
                 int64_t ldb, double* C, int64_t ldc);

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

// G (p x p, symmetric positive definite) = L L^T;  Rinv = (L^T)^-1 (upper, column-major p x p).
// Single CTA.  status[0] |= 1 when a pivot is not safely positive (caller falls back).
__global__ void __launch_bounds__(256)
chol_inv_kernel(const double* __restrict__ G, int p, double* __restrict__ Rinv, int* __restrict__ status) {
    extern __shared__ double sm[];
    const int ld = p + 1;
    double* A = sm;                                    // [p][ld], lower triangle used
    __shared__ double s_piv;
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    double tr = 0.0;
    for (int e = tid; e < p * p; e += 256) {
        const int i = e / p, j = e % p;
        A[i * ld + j] = G[i + (size_t)p * j];
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int i = 0; i < p; i++) tr = fmax(tr, A[i * ld + i]);
    const double tiny = 1e-13 * tr;                    // kappa(G) beyond ~1e13: CholeskyQR no longer trustworthy
    for (int j = 0; j < p; j++) {
        if (tid == 0) {
            const double d = A[j * ld + j];
            if (!(d > tiny)) s_bad = 1;
            s_piv = sqrt(fmax(d, tiny));
        }
        __syncthreads();
        const double ljj = s_piv;
        for (int i = j + tid; i < p; i += 256) A[i * ld + j] = (i == j) ? ljj : A[i * ld + j] / ljj;
        __syncthreads();
        // trailing update of the lower triangle: A[i][l] -= A[i][j] * A[l][j],  j < l <= i
        const int nt = p - j - 1;
        for (int e = tid; e < nt * nt; e += 256) {
            const int i = j + 1 + e / nt, l = j + 1 + e % nt;
            if (l <= i) A[i * ld + l] -= A[i * ld + j] * A[l * ld + j];
        }
        __syncthreads();
    }
    // X = L^-1 (lower); thread j owns column j and keeps it in the unused upper triangle: X[i][j] -> A[j][i]
    // (i > j); the diagonal 1/l_jj goes to xd[].  Rinv(l, i) = X(i, l): Rinv[j + p*i] = X[i][j].
    double* xd = sm + (size_t)p * ld;
    for (int j = tid; j < p; j += 256) {
        const double xjj = 1.0 / A[j * ld + j];
        xd[j] = xjj;
        for (int i = j + 1; i < p; i++) {
            double s = A[i * ld + j] * xjj;
            for (int l = j + 1; l < i; l++) s += A[i * ld + l] * A[j * ld + l];
            A[j * ld + i] = -s / A[i * ld + i];
        }
    }
    __syncthreads();
    for (int e = tid; e < p * p; e += 256) {
        const int jj = e % p, i = e / p;                       // Rinv[jj + p*i]
        Rinv[e] = (i > jj) ? A[jj * ld + i] : (i == jj ? xd[jj] : 0.0);
    }
    if (tid == 0 && s_bad) atomicOr(status, 1);
}

// Eigen-decomposition of a symmetric p x p matrix (p even) by cyclic two-sided Jacobi in shared memory.
// p/2 disjoint rotations per round, 8 threads each.  W: column-major eigenvectors, ev: eigenvalues (unsorted).
__global__ void __launch_bounds__(1024)
sym_eig_kernel(const double* __restrict__ H, int p, double* __restrict__ W, double* __restrict__ ev,
               int* __restrict__ status) {
    extern __shared__ double sm[];
    const int ld = p + 1;
    double* A = sm;
    double* V = sm + (size_t)p * ld;
    __shared__ int any_rot;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int e = tid; e < p * p; e += nthr) {
        const int i = e / p, j = e % p;
        A[i * ld + j] = 0.5 * (H[i + (size_t)p * j] + H[j + (size_t)p * i]);
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    double tr = 0.0;
    for (int i = 0; i < p; i++) tr += A[i * ld + i];
    const double floor_abs = 1e-17 * tr;
    const int k = tid >> 3, l8 = tid & 7;
    const bool active = k < p / 2;
    int sweep = 0;
    for (; sweep < 60; sweep++) {
        if (tid == 0) any_rot = 0;
        __syncthreads();
        for (int rd = 0; rd < p - 1; rd++) {
            int pp = 0, qq = 1;
            double c = 1.0, s = 0.0;
            if (active) {
                int a, b;
                if (k == 0) { a = p - 1; b = rd; }
                else { a = (rd + k) % (p - 1); b = (rd - k + (p - 1)) % (p - 1); }
                pp = min(a, b); qq = max(a, b);
                const double app = A[pp * ld + pp], aqq = A[qq * ld + qq], apq = A[pp * ld + qq];
                const double aa = fabs(apq);
                if (aa > floor_abs && aa * aa > 1e-30 * fabs(app * aqq)) {
                    const double zeta = aqq - app, beta = 2.0 * apq;
                    const double t = (zeta >= 0.0 ? beta : -beta) / (fabs(zeta) + sqrt(zeta * zeta + beta * beta));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                    if (l8 == 0) any_rot = 1;
                }
            }
            __syncwarp();
            if (active) {
                for (int r = l8; r < p; r += 8) {
                    const double x = A[r * ld + pp], y = A[r * ld + qq];
                    A[r * ld + pp] = c * x - s * y;
                    A[r * ld + qq] = s * x + c * y;
                    const double vx = V[r * ld + pp], vy = V[r * ld + qq];
                    V[r * ld + pp] = c * vx - s * vy;
                    V[r * ld + qq] = s * vx + c * vy;
                }
            }
            __syncthreads();
            if (active) {
                for (int cc = l8; cc < p; cc += 8) {
                    const double x = A[pp * ld + cc], y = A[qq * ld + cc];
                    A[pp * ld + cc] = c * x - s * y;
                    A[qq * ld + cc] = s * x + c * y;
                }
            }
            __syncthreads();
        }
        if (!any_rot) break;
        __syncthreads();
    }
    if (tid == 0 && sweep >= 60) atomicOr(status, 2);
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
}  // namespace

// M: column-major m x n (leading dimension ldm) on the device, already scaled.  trace_dev: ||M||_F^2 on the
// device or nullptr (== 1).  On success *done = true and the cores are written; *done = false means
// "not applicable / not converged": the caller must run the full Jacobi SVD.
int svd_subspace_device(mpst_ctx* c, const double* M, int64_t ldm, int m, int n, int C, int chi_max, double cutoff,
                        const double* trace_dev, double* label_core, double* ortho_core, int* chi_new,
                        double* sigma_host, bool* done) {
    *done = false;
    const int k = chi_max;
    // subspace dimension: 2k+16, capped at 112 (two p x (p+1) matrices must fit the 227 KB of shared memory of
    // the single-CTA Rayleigh-Ritz eigen-solver); at least 32 vectors of oversampling or the path is not taken
    const int p = std::min((int)round_up(2 * k + 16, 16), 112);
    if (getenv("MPST_SVD_NOSUB")) return MPST_OK;
    if (p < k + 32 || n < p + 32 || m < p) return MPST_OK;            // small / wide problems: full Jacobi
    const size_t need = (size_t)2 * n * p + (size_t)2 * m * p + 4 * (size_t)p * p + (size_t)n * k + (size_t)m * k + 4 * p + 64;
    TRY(ensure_buf(c, &c->sub, &c->subcap, need));
    double* Qa = c->sub;
    double* Qb = Qa + (size_t)n * p;
    double* Za = Qb + (size_t)n * p;
    double* Zb = Za + (size_t)m * p;
    double* Gm = Zb + (size_t)m * p;
    double* Ri = Gm + (size_t)p * p;
    double* Wm = Ri + (size_t)p * p;
    double* Wk = Wm + (size_t)p * p;
    double* T2 = Wk + (size_t)p * p;                                   // n x k
    double* Uk = T2 + (size_t)n * k;                                   // m x k
    double* ev = Uk + (size_t)m * k;                                   // p, then Psorted p
    int* status = c->iscal + 8;
    unsigned long long* resbits = reinterpret_cast<unsigned long long*>(c->scal + 10);
    const size_t chol_smem = sizeof(double) * ((size_t)p * (p + 1) + p);
    const size_t eig_smem = 2 * sizeof(double) * (size_t)p * (p + 1);
    CUDA_TRY(c, cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem));
    CUDA_TRY(c, cudaFuncSetAttribute(sym_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig_smem));
    CUDA_TRY(c, cudaMemsetAsync(status, 0, sizeof(int), c->stream));

    // X (rows x p, ld = rows) -> orthonormal columns in `out`; `tmp` is scratch of the same size
    auto cholqr = [&](double* X, double* out, int rows) -> int {
        TRY(launch_dgemm(c, 1, 0, p, p, rows, X, rows, X, rows, Gm, p));
        chol_inv_kernel<<<1, 256, chol_smem, c->stream>>>(Gm, p, Ri, status);
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
        const int niter = round == 0 ? (p >= 2 * k ? 4 : 6) : 3;
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
        TRY(launch_dgemm(c, 0, 0, m, p, n, M, ldm, Qa, n, Za, m));                // Z = M Q
        TRY(launch_dgemm(c, 1, 0, p, p, m, Za, m, Za, m, Gm, p));                 // H = Z^T Z
        sym_eig_kernel<<<1, (unsigned)round_up((p / 2) * 8, 32), eig_smem, c->stream>>>(Gm, p, Wm, ev, status);
        ritz_trunc_kernel<<<1, 256, 0, c->stream>>>(ev, p, n, trace_dev, chi_max, cutoff, c->perm, ev + p, c->iscal);
        gather_cols_kernel<<<(p * k + 255) / 256, 256, 0, c->stream>>>(Wm, p, c->perm, k, Wk);
        c->launches += 3;
        TRY(launch_dgemm(c, 0, 0, n, k, p, Qa, n, Wk, p, ortho_core, n));         // V_k = Q W_k  (all k columns; chi_new <= k used)
        TRY(launch_dgemm(c, 0, 0, m, k, p, Za, m, Wk, p, Uk, m));                 // U_k S_k = Z W_k
        // residual of the kept Ritz pairs: M^T (M v_i) - sigma_i^2 v_i
        TRY(launch_dgemm(c, 1, 0, n, k, m, M, ldm, Uk, m, T2, n));
        CUDA_TRY(c, cudaMemsetAsync(resbits, 0, sizeof(unsigned long long), c->stream));
        residual_kernel<<<k, 256, 0, c->stream>>>(T2, ortho_core, ev + p, c->iscal, n, resbits);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal, c->iscal, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal + 8, status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->hscal + 10, resbits, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        const double res = c->hscal[10];
        if (getenv("MPST_SVD_DEBUG"))
            fprintf(stderr, "[svd subspace] m=%d n=%d p=%d iters=%d chi=%d status=%d residual=%.2e\n", m, n, p, iters_done,
                    c->hiscal[0], c->hiscal[8], res);
        if (c->hiscal[8] != 0 || !(res == res)) return MPST_OK;                    // breakdown: full Jacobi
        if (res > (round == 0 ? 1e-5 : 1e-10)) return MPST_OK;                     // spectrum too flat: full Jacobi
        if (res <= 5e-14) {
            *chi_new = c->hiscal[0];
            scatter_label_kernel<<<2 * c->sm_count, 256, 0, c->stream>>>(Uk, m, m / C, c->iscal, label_core);
            c->launches++;
            CUDA_TRY(c, cudaGetLastError());
            if (sigma_host) {
                std::vector<double> tmp(*chi_new);
                CUDA_TRY(c, cudaMemcpy(tmp.data(), ev + p, sizeof(double) * (*chi_new), cudaMemcpyDeviceToHost));
                for (int q = 0; q < *chi_new; q++) sigma_host[q] = sqrt(tmp[q]);
            }
            *done = true;
            return MPST_OK;
        }
    }
    return MPST_OK;                                                                // not converged: full Jacobi
}
