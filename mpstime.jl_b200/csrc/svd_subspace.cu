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
#include <cstring>
#include <utility>
#include "mpst_common.cuh"

int launch_dgemm(mpst_ctx* c, int ta, int tb, int M, int N, int K, const double* A, int64_t lda, const double* B,
                 int64_t ldb, double* C, int64_t ldc);
int launch_dgemm_splitk(mpst_ctx* c, int ta, int tb, int M, int N, int K, const double* A, int64_t lda, const double* B,
                        int64_t ldb, double* Cparts, int max_splits, int* splits_out);
int launch_dgemm_on(mpst_ctx* c, int lane, int ta, int tb, int M, int N, int K, const double* A, int64_t lda,
                    const double* B, int64_t ldb, double* C, int64_t ldc);
bool launch_sym_eig_reg(int p, const double* L, double* W, double* ev, int* status, cudaStream_t st);

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
// Rinv = (L^T)^-1 = (L^-1)^T (upper, column-major p x p), p <= 16 NR.
// Single CTA of 16 x 16 threads; the whole matrix lives in REGISTERS (thread (ty, tx) owns the NR x NR elements
// (ty + 16 r, tx + 16 c)), so a column step is one shared-memory broadcast of row k and one block barrier, and
// the instruction stream per step is ~NR^2 DFMAs plus a short scalar preamble (these kernels are issue-bound,
// not flop-bound: 8 warps instead of 32 is what makes the step short).
// Right-looking Cholesky fused with the forward substitution on the identity, in place: after step k the
// positions (i, j <= k) that held L are dead and take the running inverse X = L^-1, the trailing square
// (i, j > k) keeps the (full, symmetric) Schur complement, so row k alone carries everything step k needs:
//     inv = 1/sqrt(m_kk);  l_i = m_ki inv (i > k);  r_j = m_kj inv
//     m_ij -= l_i r_j (i > k, j != k);  m_ik = -l_i inv (i > k);  m_kj = r_j (j < k);  m_kk = inv.
// status[0] |= 1 on NaN input (caller falls back to the exact SVD); numerically dependent columns are deflated (below).
template <int NRR, int NRC, int TYN>
__global__ void __launch_bounds__(TYN * 16)
chol_inv_kernel(const double* __restrict__ G, int splits, int p, double* __restrict__ Rinv, double* __restrict__ Lout,
                int* __restrict__ status) {
    __shared__ double rowk[2][128];
    __shared__ double s_diag[128];
    constexpr int NTHR = TYN * 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double m[NRR][NRC];
#ifdef MPST_KDEBUG
    const long long tk00 = clock64();
#endif
    for (int t = threadIdx.x; t < 256; t += NTHR) (&rowk[0][0])[t] = 0.0;          // tails [p, 128) stay zero
    // sum the split-K partials (fixed order, partial-major so a thread has all its loads of one partial in flight).
    // The Gram partials are symmetric (a*b == b*a, same accumulation order in both triangles), so the transposed,
    // coalesced read is the same matrix.
#pragma unroll
    for (int r = 0; r < NRR; r++)
#pragma unroll
        for (int cc = 0; cc < NRC; cc++) m[r][cc] = 0.0;
    int roff[NRR], coff[NRC];                          // clamped -> unconditional loads that batch (one latency per partial)
#pragma unroll
    for (int r = 0; r < NRR; r++) roff[r] = p * min(ty + TYN * r, p - 1);
#pragma unroll
    for (int cc = 0; cc < NRC; cc++) coff[cc] = min(tx + 16 * cc, p - 1);
    for (int z = 0; z < splits; z++) {
        const double* gz = G + (size_t)z * p * p;
        double v[NRR][NRC];
#pragma unroll
        for (int r = 0; r < NRR; r++)
#pragma unroll
            for (int cc = 0; cc < NRC; cc++) v[r][cc] = gz[roff[r] + coff[cc]];    // (j, i): coalesced along tx
#pragma unroll
        for (int r = 0; r < NRR; r++)
#pragma unroll
            for (int cc = 0; cc < NRC; cc++) m[r][cc] += v[r][cc];
    }
#pragma unroll
    for (int r = 0; r < NRR; r++)
#pragma unroll
        for (int cc = 0; cc < NRC; cc++)
            if (ty + TYN * r >= p || tx + 16 * cc >= p) m[r][cc] = 0.0;
#pragma unroll
    for (int r = 0; r < NRR; r++)
#pragma unroll
        for (int cc = 0; cc < NRC; cc++)
            if (ty + TYN * r == tx + 16 * cc && ty + TYN * r < p) s_diag[ty + TYN * r] = m[r][cc];
    __syncthreads();
    // A pivot is the squared norm of column k of the factored matrix after projecting out columns < k; relative to
    // the column's own squared norm g_kk it is sin^2 of the angle to their span.  The test is per column (Cholesky
    // itself is invariant under column scaling), so graded columns -- norms spread over orders of magnitude but
    // directions well separated -- pass.  A column below 1e-13 is numerically inside the span of its predecessors
    // (the tail of a bond tensor below rounding level, or a rank-deficient bond early in training): it is DEFLATED --
    // 1/l_kk := 0 makes row and column k of both L and L^-1 zero, so X R^-1 simply has a zero column there and every
    // other column is what it would be without column k.  A zero column stays zero through M, M^T, the second pass and
    // the Rayleigh-Ritz step (Ritz value 0, ranked last).  status bit 0 is left for NaN input.
    bool bad = false;
#ifdef MPST_KDEBUG
    const long long tk0 = clock64();
#endif
    // The pivot loop is unrolled over the (row slot, column slot) that holds the pivot, so every register index is
    // a compile-time constant (no jump tables) and the finished row slots above the pivot drop out of the update.
    constexpr int HS = 16 / TYN;                       // row slots per 16-column block (2 for TYN = 8)
#pragma unroll
    for (int kb = 0; kb < NRC; kb++) {
#pragma unroll
        for (int h = 0; h < HS; h++) {
            const int krr = kb * HS + h;               // row slot of the pivots k0 .. k0 + TYN - 1
            const int k0 = 16 * kb + TYN * h;
            const int kend = min(TYN, p - k0);
#pragma unroll 1
            for (int kt = 0; kt < kend; kt++) {
                const int k = k0 + kt;
                double* rk = &rowk[k & 1][0];
                if (ty == kt) {                        // the 16 threads that own row k publish it
#pragma unroll
                    for (int cc = 0; cc < NRC; cc++) rk[tx + 16 * cc] = m[krr][cc];   // columns >= p hold zeros
                }
                __syncthreads();                       // the only barrier of the step (row buffers alternate)
                const double dk = rk[k];
                const double tiny = 1e-13 * s_diag[k];
                bad |= !(dk == dk);
                const double inv = dk > tiny ? rsqrt(dk) : 0.0;       // 1 / l_kk, or 0: column k is deflated
                double li[NRR], rj[NRC];
#pragma unroll
                for (int r = krr; r < NRR; r++) {
                    const double v = rk[ty + TYN * r] * inv;          // every index < 128 is initialised
                    li[r] = (r > krr || ty > kt) ? v : 0.0;           // rows at or above the pivot do not move
                }
#pragma unroll
                for (int cc = 0; cc < NRC; cc++) rj[cc] = rk[tx + 16 * cc] * inv;
#pragma unroll
                for (int r = krr; r < NRR; r++)
#pragma unroll
                    for (int cc = 0; cc < NRC; cc++) m[r][cc] = fma(-li[r], rj[cc], m[r][cc]);
                const bool col_owner = tx == TYN * h + kt;            // column k below the pivot: dead L -> X[i][k]
                if (Lout != nullptr && col_owner) {                   // column k of L (diagonal = sqrt(d_k)), rows >= slot start
#pragma unroll
                    for (int r = krr; r < NRR; r++) {
                        const int i = ty + TYN * r;
                        if (i < p) Lout[i + (size_t)p * k] = (r == krr && ty == kt) ? dk * inv : li[r];
                    }
                }
#pragma unroll
                for (int r = krr; r < NRR; r++) m[r][kb] = (col_owner && (r > krr || ty > kt)) ? -li[r] * inv : m[r][kb];
                if (ty == kt) {                        // row k: finished row of X (columns > k are dead)
#pragma unroll
                    for (int cc = 0; cc < NRC; cc++) m[krr][cc] = (cc == kb && col_owner) ? inv : rj[cc];
                }
            }
        }
    }
#ifdef MPST_KDEBUG
    if (threadIdx.x == 0) printf("[chol] p=%d splits=%d loop cycles %lld  since start %lld\n", p, splits, clock64() - tk0, clock64() - tk00);
#endif
#pragma unroll
    for (int r = 0; r < NRR; r++)
#pragma unroll
        for (int cc = 0; cc < NRC; cc++) {
            const int i = ty + TYN * r, j = tx + 16 * cc;                 // Rinv[j + p*i] = X[i][j]  (i >= j)
            if (i < p && j < p) Rinv[j + (size_t)p * i] = (i >= j) ? m[r][cc] : 0.0;
        }
    if (threadIdx.x == 0 && bad) atomicOr(status, 1);
}

// Blocked form of chol_inv_kernel for p = 16 NRC: panels of 8 pivots.  The unblocked kernel pays one block barrier
// and one dependent rsqrt chain per pivot (47 us at p = 112: ~825 cycles per pivot, 5x the FP64 work).  Here one warp
// factors and inverts the 8 x 8 diagonal block in registers (the only serial part), 16 NRC threads form the panel
//     P[:, x] = Lkk^-1 M[K, x]   (x outside the panel: column x of L^T below the panel / finished rows of X left of it)
//     P[:, K] = Lkk^-1
// and every thread applies the eight rank-1 updates m_ij -= P[t][i] P[t][j] from shared memory with no barrier in
// between: three barriers per 8 pivots.  Same in-place layout as the unblocked kernel (trailing square = symmetric
// Schur complement, dead L positions take X = L^-1), same per-column breakdown test.
template <int NRC>
__global__ void __launch_bounds__(256)
chol_inv_blk_kernel(const double* __restrict__ G, double* __restrict__ Rinv, double* __restrict__ Lout,
                    int* __restrict__ status) {
    constexpr int P = 16 * NRC, PP = P + 4;
    __shared__ double praw[8][PP];
    __shared__ double pm[8][PP];
    __shared__ __align__(16) double linv[8][8];
    __shared__ double lkk[8][8];
    __shared__ double s_diag[P];
    __shared__ int s_bad;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double m[NRC][NRC];
    if (threadIdx.x == 0) s_bad = 0;
    // G is symmetric (same accumulation order in both triangles): the transposed, coalesced read is the same matrix
#pragma unroll
    for (int r = 0; r < NRC; r++)
#pragma unroll
        for (int cc = 0; cc < NRC; cc++) m[r][cc] = G[(size_t)P * (ty + 16 * r) + tx + 16 * cc];
#pragma unroll
    for (int r = 0; r < NRC; r++)
#pragma unroll
        for (int cc = 0; cc < NRC; cc++)
            if (r == cc && ty == tx) s_diag[ty + 16 * r] = m[r][cc];
    __syncthreads();
#pragma unroll
    for (int kp = 0; kp < 2 * NRC; kp++) {
        const int k0 = 8 * kp, slot = kp >> 1, half = kp & 1;
        const bool in_panel = (ty >> 3) == half;                 // this thread's row of slot `slot` is a pivot row
        // (a) publish the 8 pivot rows (already updated by the previous panels)
        if (in_panel) {
#pragma unroll
            for (int cc = 0; cc < NRC; cc++) praw[ty & 7][tx + 16 * cc] = m[slot][cc];
        }
        __syncthreads();
        // (b) warp 0: Cholesky of the 8 x 8 diagonal block and its inverse, all lanes redundantly in registers
        if (threadIdx.x < 32) {
            double a[8][8], inv[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) a[i][j] = praw[i][k0 + j];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double dk = a[k][k];
                const double tiny = 1e-13 * s_diag[k0 + k];
                bad |= !(dk == dk);
                inv[k] = dk > tiny ? rsqrt(dk) : 0.0;             // 0: column deflated (see chol_inv_kernel)
                a[k][k] = dk * inv[k];
#pragma unroll
                for (int i = k + 1; i < 8; i++) a[i][k] *= inv[k];
#pragma unroll
                for (int j = k + 1; j < 8; j++)
#pragma unroll
                    for (int i = j; i < 8; i++) a[i][j] = fma(-a[i][k], a[j][k], a[i][j]);
            }
            // inverse: lane k (< 8) forms column k of X = Lkk^-1 by forward substitution (x_j = 0 for j < k), with
            // predicates instead of lane-dependent loop bounds so that every register index stays a constant
            const int lane = threadIdx.x;
            const int kc = lane & 7;
            double xc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                double sacc = 0.0;
#pragma unroll
                for (int j = 0; j < i; j++) sacc = fma(a[i][j], xc[j], sacc);
                xc[i] = i < kc ? 0.0 : (i == kc ? inv[i] : -sacc * inv[i]);
            }
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < 8; i++) linv[i][kc] = xc[i];
            }
            if (lane == 8 && Lout != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) lkk[i][j] = j <= i ? a[i][j] : 0.0;
            }
            if (lane == 0 && bad) s_bad = 1;
        }
        __syncthreads();
        // (c) the panel: one thread per column x
        if (threadIdx.x < P) {
            const int x = threadIdx.x;
            const bool diag = x >= k0 && x < k0 + 8;
            double v[8], o[8];
#pragma unroll
            for (int s = 0; s < 8; s++) v[s] = praw[s][x];
#pragma unroll
            for (int t = 0; t < 8; t++) {
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s <= t; s++) acc = fma(linv[t][s], v[s], acc);
                o[t] = diag ? linv[t][(x - k0) & 7] : acc;
                pm[t][x] = o[t];
            }
            if (Lout != nullptr) {
                if (x >= k0 + 8) {
#pragma unroll
                    for (int t = 0; t < 8; t++) Lout[x + (size_t)P * (k0 + t)] = o[t];
                } else if (diag) {
#pragma unroll
                    for (int t = 0; t < 8; t++)
                        if (t <= x - k0) Lout[x + (size_t)P * (k0 + t)] = lkk[x - k0][t];
                }
            }
        }
        __syncthreads();
        // (d) rows of the panel take P; rows below take the eight rank-1 updates (columns of the panel: X = -L_iK Lkk^-1)
        const bool cpan = (tx >> 3) == half;                     // this thread's column of slot `slot` is a panel column
#pragma unroll
        for (int r = slot; r < NRC; r++) {
            if (r == slot && in_panel) {
#pragma unroll
                for (int cc = 0; cc < NRC; cc++) m[r][cc] = pm[ty & 7][tx + 16 * cc];
            } else if (cpan && (r > slot || (ty >> 3) > half)) {
                m[r][slot] = 0.0;                                // consumed L entries: start of X_iK
            }
        }
#pragma unroll 2
        for (int t = 0; t < 8; t++) {
            double li[NRC], rj[NRC];
#pragma unroll
            for (int r = slot; r < NRC; r++) {
                const double vv = pm[t][ty + 16 * r];
                li[r] = (r > slot || (ty >> 3) > half) ? vv : 0.0;       // rows at or above the panel do not move
            }
#pragma unroll
            for (int cc = 0; cc < NRC; cc++) rj[cc] = pm[t][tx + 16 * cc];
#pragma unroll
            for (int r = slot; r < NRC; r++)
#pragma unroll
                for (int cc = 0; cc < NRC; cc++) m[r][cc] = fma(-li[r], rj[cc], m[r][cc]);
        }
    }
#pragma unroll
    for (int r = 0; r < NRC; r++)
#pragma unroll
        for (int cc = 0; cc < NRC; cc++) {
            const int i = ty + 16 * r, j = tx + 16 * cc;                  // Rinv[j + p*i] = X[i][j]  (i >= j)
            Rinv[j + (size_t)P * i] = (i >= j) ? m[r][cc] : 0.0;
        }
    __syncthreads();
    if (threadIdx.x == 0 && s_bad) atomicOr(status, 1);
}

// p <= 96: 4 warps (8 x 16 threads, 2NC x NC register tile) -- these kernels are issue/latency bound and the
// per-thread scalar preamble dominates, so fewer, fatter threads win; larger p: 8 warps (register budget).
void launch_chol_inv(int p, const double* G, int splits, double* Rinv, double* Lout, int* status, cudaStream_t st,
                     bool blocked = true) {
    if (blocked && splits == 1 && p % 16 == 0 && p >= 32 && p <= 128) {
        switch (p / 16) {
            case 2: chol_inv_blk_kernel<2><<<1, 256, 0, st>>>(G, Rinv, Lout, status); return;
            case 3: chol_inv_blk_kernel<3><<<1, 256, 0, st>>>(G, Rinv, Lout, status); return;
            case 4: chol_inv_blk_kernel<4><<<1, 256, 0, st>>>(G, Rinv, Lout, status); return;
            case 5: chol_inv_blk_kernel<5><<<1, 256, 0, st>>>(G, Rinv, Lout, status); return;
            case 6: chol_inv_blk_kernel<6><<<1, 256, 0, st>>>(G, Rinv, Lout, status); return;
            case 7: chol_inv_blk_kernel<7><<<1, 256, 0, st>>>(G, Rinv, Lout, status); return;
            default: chol_inv_blk_kernel<8><<<1, 256, 0, st>>>(G, Rinv, Lout, status); return;
        }
    }
    switch ((p + 15) / 16) {
        case 1: chol_inv_kernel<2, 1, 8><<<1, 128, 0, st>>>(G, splits, p, Rinv, Lout, status); break;
        case 2: chol_inv_kernel<4, 2, 8><<<1, 128, 0, st>>>(G, splits, p, Rinv, Lout, status); break;
        case 3: chol_inv_kernel<6, 3, 8><<<1, 128, 0, st>>>(G, splits, p, Rinv, Lout, status); break;
        case 4: chol_inv_kernel<8, 4, 8><<<1, 128, 0, st>>>(G, splits, p, Rinv, Lout, status); break;
        case 5: chol_inv_kernel<10, 5, 8><<<1, 128, 0, st>>>(G, splits, p, Rinv, Lout, status); break;
        case 6: chol_inv_kernel<12, 6, 8><<<1, 128, 0, st>>>(G, splits, p, Rinv, Lout, status); break;
        case 7: chol_inv_kernel<7, 7, 16><<<1, 256, 0, st>>>(G, splits, p, Rinv, Lout, status); break;
        default: chol_inv_kernel<8, 8, 16><<<1, 256, 0, st>>>(G, splits, p, Rinv, Lout, status); break;
    }
}

// Eigen-decomposition of a symmetric positive semi-definite q x q matrix (a Gram matrix; embedded in an even
// order p >= q, p <= 16 NT <= 128) by ONE-SIDED (Hestenes) Jacobi on its own columns, one CTA: right rotations J
// make the columns of H J mutually orthogonal, then H J = W Lambda, i.e. J = W and lambda_i = ||(H J)_i||.
// A half-warp owns one column pair per round (round-robin ordering, p/2 disjoint pairs, 64 half-warps): it pulls
// its two columns of H J into registers, forms alpha = |h_a|^2, beta = |h_b|^2, gamma = h_a.h_b with 4 shuffle
// steps, rotates them and the two columns of J and writes back -- one pass over both matrices and ONE barrier
// per round; rotation from two rsqrt, no division or sqrt (the kernel is instruction-issue bound).
// H carries absolute errors ~eps*trace per entry, so gamma is noise below ~eps*trace*max(|h_a|,|h_b|): that
// is the rotation floor (rotating noise never terminates; the eigenvalues are resolved to eps*trace anyway).
// W: column-major eigenvectors, ev: eigenvalues (unsorted).  status[1] = sweeps used.
template <int NT, bool FULL, bool LMODE>
__global__ void __launch_bounds__(1024)
sym_eig_kernel(const double* __restrict__ H, int splits, int q, int p, double* __restrict__ W, double* __restrict__ ev,
               int* __restrict__ status) {
    extern __shared__ double sm[];
    double* Hc = sm;                                   // [p][p] columns of H J
    double* Vc = sm + (size_t)p * p;                   // [p][p] columns of J
    double* nrm = Vc + (size_t)p * p;                  // [p] cached squared column norms of H J
    __shared__ int any_rot, any_big;
#ifdef MPST_KDEBUG
    __shared__ int dbg_rot, dbg_big;
    __shared__ unsigned long long dbg_max;
    if (threadIdx.x == 0) { dbg_rot = 0; dbg_big = 0; dbg_max = 0ull; }
#endif
    const int tid = threadIdx.x, l16 = tid & 15, hw = tid >> 4, hp = p / 2;
    double tr = 0.0;
    if (LMODE) {
        // H is the lower Cholesky factor L of the matrix to decompose (L L^T = W Lambda W^T): Jacobi on the columns of
        // L gives L J = W Sigma, no rotation accumulator needed; trace(L L^T) = ||L||_F^2
        for (int e = tid; e < p * p; e += 1024) {
            const int i = e % p, j = e / p;
            Hc[e] = (i >= j) ? H[e] : 0.0;
        }
        __syncthreads();
        for (int j = hw; j < p; j += 64) {
            double nn = 0.0;
            for (int l = l16; l < p; l += 16) nn = fma(Hc[j * p + l], Hc[j * p + l], nn);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o, 16);
            if (l16 == 0) nrm[j] = nn;
        }
        __syncthreads();
        for (int j = 0; j < p; j++) tr += nrm[j];
        __syncthreads();
    } else {
        // sum the split-K partials with coalesced reads (fixed order per element), then symmetrise in shared memory
        for (int e = tid; e < p * p; e += 1024) {
            const int i = e % p, j = e / p;
            double v = 0.0;
            if (i < q && j < q)
                for (int z = 0; z < splits; z++) v += H[(size_t)z * q * q + i + (size_t)q * j];
            Hc[e] = v;
            Vc[e] = (i == j) ? 1.0 : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < p * p; e += 1024) {
            const int i = e % p, j = e / p;
            if (i < j) {
                const double v = 0.5 * (Hc[i + p * j] + Hc[j + p * i]);
                Hc[i + p * j] = v;
                Hc[j + p * i] = v;
            }
        }
        __syncthreads();
        for (int i = 0; i < p; i++) tr += Hc[i * p + i];
    }
    // rotation floor on gamma^2: columns of H J carry absolute noise eps*trace, columns of L J carry eps*sqrt(trace)
    const double thr2 = LMODE ? (2.2e-16 * 2.2e-16) * tr : (8.9e-16 * tr) * (8.9e-16 * tr);
    const bool active = hw < hp;
    const bool warp_active = (tid >> 5) * 2 < hp;      // warps without a pair only take part in the barriers
    int sweep = 0;
#ifdef MPST_KDEBUG
    const long long tk0 = clock64();
#endif
    for (; sweep < 60; sweep++) {
        if (tid == 0) { any_rot = 0; any_big = 0; }
        // squared column norms, recomputed from the columns once per sweep and updated by the rotations in between
        // (they only steer the rotation; the eigenvalues are taken from the columns themselves at the end)
        for (int j = hw; j < p; j += 64) {
            double nn = 0.0;
            for (int l = l16; l < p; l += 16) nn = fma(Hc[j * p + l], Hc[j * p + l], nn);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o, 16);
            if (l16 == 0) nrm[j] = nn;
        }
        __syncthreads();
        // round-robin: half-warp 0 pairs (p-1, rd); half-warp h pairs ((rd+h) mod (p-1), (rd-h) mod (p-1))
        int a = (hw == 0) ? p - 1 : hw, b = (hw == 0) ? 0 : p - 1 - hw;
        for (int rd = 0; rd < p - 1; rd++) {
            if (warp_active) {
                const int ca = active ? min(a, b) : 0, cb = active ? max(a, b) : 1;
                double* ha = Hc + ca * p + l16;
                double* hb = Hc + cb * p + l16;
                double xa[NT], xb[NT];
                double ga = 0.0;
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const bool in = active && (FULL || (l16 + 16 * t < p));   // a half-warp without a pair reads nothing
                    xa[t] = in ? ha[16 * t] : 0.0;
                    xb[t] = in ? hb[16 * t] : 0.0;
                    ga = fma(xa[t], xb[t], ga);
                }
                const double al = active ? nrm[ca] : 0.0, be = active ? nrm[cb] : 0.0;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) ga += __shfl_xor_sync(0xffffffffu, ga, o, 16);
                const double g2 = ga * ga, ab = al * be;
                if (active && g2 > thr2 * fmax(al, be) && g2 > 1e-30 * ab) {
                    // cos/sin of the double angle, then half-angle: two rsqrt
                    const double zeta = be - al, beta2 = 2.0 * ga;
                    const double inv_r = rsqrt(zeta * zeta + beta2 * beta2);
                    const double cos2 = fabs(zeta) * inv_r, sin2 = (zeta >= 0.0 ? beta2 : -beta2) * inv_r;
                    const double c2 = 0.5 + 0.5 * cos2;
                    const double inv_c = rsqrt(c2);
                    const double c = c2 * inv_c, sn = 0.5 * sin2 * inv_c;
                    double* va = Vc + ca * p + l16;
                    double* vb = Vc + cb * p + l16;
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        if (FULL || (l16 + 16 * t < p)) {
                            ha[16 * t] = c * xa[t] - sn * xb[t];
                            hb[16 * t] = sn * xa[t] + c * xb[t];
                            if (!LMODE) {
                                const double ya = va[16 * t], yb = vb[16 * t];
                                va[16 * t] = c * ya - sn * yb;
                                vb[16 * t] = sn * ya + c * yb;
                            }
                        }
                    }
#ifdef MPST_KDEBUG
                    if (l16 == 0) {
                        atomicAdd(&dbg_rot, 1);
                        if (g2 > 1e-16 * ab) atomicAdd(&dbg_big, 1);
                        atomicMax(&dbg_max, (unsigned long long)__double_as_longlong(g2 / ab));
                    }
#endif
                    if (l16 == 0) {
                        // |h_a'|^2 = c^2 al - 2cs ga + s^2 be,  |h_b'|^2 = s^2 al + 2cs ga + c^2 be
                        const double cs2 = 2.0 * c * sn * ga, cc = c * c, ss = sn * sn;
                        nrm[ca] = cc * al - cs2 + ss * be;
                        nrm[cb] = ss * al + cs2 + cc * be;
                        atomicOr(&any_rot, 1);                    // (atomics: several half-warps raise the same flag)
                        if (g2 > 1e-16 * ab) atomicOr(&any_big, 1);   // cosine between the columns above 1e-8
                    }
                }
            }
            if (hw != 0) a = (a + 1 == p - 1) ? 0 : a + 1;
            b = (hw == 0) ? b + 1 : ((b + 1 == p - 1) ? 0 : b + 1);
            __syncthreads();
        }
#ifdef MPST_KDEBUG
        if (tid == 0) {
            printf("[eig sweep %d] rotations %d big %d max cos %.3e\n", sweep, dbg_rot, dbg_big, sqrt(__longlong_as_double((long long)dbg_max)));
            dbg_rot = 0; dbg_big = 0; dbg_max = 0ull;
        }
#endif
        // quadratic convergence: a sweep whose largest rotated cosine was <= 1e-8 leaves cosines at ~1e-16
        if (!any_rot || !any_big) break;
        __syncthreads();
    }
#ifdef MPST_KDEBUG
    if (tid == 0) printf("[eig] p=%d q=%d splits=%d sweeps=%d loop cycles %lld\n", p, q, splits, sweep, clock64() - tk0);
#endif
    if (tid == 0 && sweep >= 60) atomicOr(status, 2);
    if (tid == 0) status[1] = sweep;
    if (!LMODE)
        for (int e = tid; e < p * p; e += 1024) W[e] = Vc[e];
    for (int j = hw; j < p; j += 64) {                 // lambda_j = || (H J)_j ||  resp.  || (L J)_j ||^2, w_j = (L J)_j / sigma_j
        double nn = 0.0;
        for (int l = l16; l < p; l += 16) nn += Hc[j * p + l] * Hc[j * p + l];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o, 16);
        if (LMODE) {
            const double inv = nn > 0.0 ? rsqrt(nn) : 0.0;
            for (int l = l16; l < p; l += 16) W[(size_t)j * p + l] = Hc[j * p + l] * inv;
            if (l16 == 0) ev[j] = nn;
        } else if (l16 == 0) ev[j] = sqrt(nn);
    }
}

template <int NT, bool LMODE>
int launch_sym_eig_nt(int p, size_t smem, const double* H, int splits, int q, double* W, double* ev, int* status, cudaStream_t st) {
    if (p % 16 == 0) {
        cudaError_t e = cudaFuncSetAttribute(sym_eig_kernel<NT, true, LMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sym_eig_kernel<NT, true, LMODE><<<1, 1024, smem, st>>>(H, splits, q, p, W, ev, status);
    } else {
        cudaError_t e = cudaFuncSetAttribute(sym_eig_kernel<NT, false, LMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sym_eig_kernel<NT, false, LMODE><<<1, 1024, smem, st>>>(H, splits, q, p, W, ev, status);
    }
    return 0;
}

// LMODE = false: H = `splits` partial Gram matrices (q x q);  LMODE = true: H = lower Cholesky factor (p x p, q == p)
template <bool LMODE>
int launch_sym_eig(int p, size_t smem, const double* H, int splits, int q, double* W, double* ev, int* status, cudaStream_t st) {
    switch ((p + 15) / 16) {
        case 1: return launch_sym_eig_nt<1, LMODE>(p, smem, H, splits, q, W, ev, status, st);
        case 2: return launch_sym_eig_nt<2, LMODE>(p, smem, H, splits, q, W, ev, status, st);
        case 3: return launch_sym_eig_nt<3, LMODE>(p, smem, H, splits, q, W, ev, status, st);
        case 4: return launch_sym_eig_nt<4, LMODE>(p, smem, H, splits, q, W, ev, status, st);
        case 5: return launch_sym_eig_nt<5, LMODE>(p, smem, H, splits, q, W, ev, status, st);
        case 6: return launch_sym_eig_nt<6, LMODE>(p, smem, H, splits, q, W, ev, status, st);
        case 7: return launch_sym_eig_nt<7, LMODE>(p, smem, H, splits, q, W, ev, status, st);
        default: return launch_sym_eig_nt<8, LMODE>(p, smem, H, splits, q, W, ev, status, st);
    }
}

// out[:, j] = in[:, j] / ||in[:, j]||  (one block per column).  Replaces the Cholesky-QR of Z = M Q in the later
// subspace iterations: with Q close to the singular vectors the columns of Z are nearly orthogonal already and only
// their norms (sigma_j) differ by orders of magnitude.
__global__ void __launch_bounds__(256)
colscale_kernel(const double* __restrict__ in, double* __restrict__ out, int rows) {
    __shared__ double sh[8];
    const double* x = in + (size_t)blockIdx.x * rows;
    double* y = out + (size_t)blockIdx.x * rows;
    double s = 0.0;
    for (int i = threadIdx.x; i < rows; i += 256) s = fma(x[i], x[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += sh[w];
    const double inv = t > 0.0 ? rsqrt(t) : 0.0;
    for (int i = threadIdx.x; i < rows; i += 256) y[i] = x[i] * inv;
}

// rank the p Ritz values, apply the NDTensors truncation rule with the weight outside the subspace
// (trace - sum) already discarded; perm[k] = column of the k-th largest, Psorted, iscal[0] = chi_new.
__global__ void __launch_bounds__(256)
ritz_trunc_kernel(const double* __restrict__ ev, int p, int n_total, const double* __restrict__ trace_dev, int maxdim,
                  double cutoff, int* __restrict__ perm, double* __restrict__ Psorted, int* __restrict__ iscal,
                  const double* __restrict__ kept_dev = nullptr) {
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
        // kept_dev (second pass of a deflated split): weight already kept by the first pass; this pass may keep nothing
        const double kept = kept_dev ? *kept_dev : 0.0;
        double err = fmax(scale - kept - sum, 0.0);        // everything outside the Ritz subspace
        (void)n_total;
        int keep = p;
        while (keep > maxdim) { err += Psorted[keep - 1]; keep--; }
        const int floor_keep = kept_dev ? 0 : 1;
        while (keep > floor_keep && err + Psorted[keep - 1] <= cutoff * scale) { err += Psorted[keep - 1]; keep--; }
        iscal[0] = max(keep, floor_keep);
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

// res = max_i || T2[:, i] - P_i * Vk[:, i] || / P_0   over the kept columns.  scale_dev (second pass of a deflated
// split): sigma_1^2 of the whole matrix from the first pass -- the deflated matrix carries the first pass's rounding
// noise (eps sigma_1 per entry), so its own leading value is the wrong yardstick.
__global__ void __launch_bounds__(256)
residual_kernel(const double* __restrict__ T2, const double* __restrict__ Vk, const double* __restrict__ Psorted,
                const int* __restrict__ iscal, int n, unsigned long long* __restrict__ out_bits,
                const double* __restrict__ scale_dev = nullptr) {
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
        const double rel = sqrt(t) / fmax(scale_dev ? *scale_dev : Psorted[0], 1e-300);
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

// ---- deflated two-pass split (chi_max > 80) ----
// out[0] = sum of the first cnt entries (fixed order)
__global__ void sum_first_kernel(const double* __restrict__ P, int cnt, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < cnt; q++) s += P[q];
        out[0] = s;
    }
}
// T <- M - T   (m x n; M with leading dimension ldm, T dense)
__global__ void __launch_bounds__(256)
deflate_kernel(const double* __restrict__ M, int64_t ldm, double* __restrict__ T, int m, int n) {
    const int64_t tot = (int64_t)m * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / m, i = e - j * m;
        T[e] = M[i + ldm * j] - T[e];
    }
}
// scatter_label_kernel with the kept dimension known on the host; also leaves it in iscal[0]
__global__ void scatter_label_n_kernel(const double* __restrict__ UkSk, int m, int Dx, int chi, int* __restrict__ iscal,
                                       double* __restrict__ label_core) {
    if (blockIdx.x == 0 && threadIdx.x == 0) iscal[0] = chi;
    const int64_t tot = (int64_t)m * chi;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e / m), r = (int)(e % m);
        const int cls = r / Dx, xx = r - cls * Dx;
        label_core[(size_t)cls * Dx * chi + xx + (size_t)Dx * kk] = UkSk[e];
    }
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
static int subspace_core(mpst_ctx* c, const double* M, int64_t ldm, int m, int n, int C, int chi_max, double cutoff,
                         const double* trace_dev, double* label_core, double* ortho_core, int* chi_new,
                         double* sigma_host, bool* done) {
    *done = false;
    c->last[L_SVD_RESTARTS] = 0;
    if (c->flag[F_SVD_NOSUB] || cutoff < 1e-12) return MPST_OK;
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
        // p = 2k, but never fewer than 32 oversampling columns (k < 32: config A's chi = 20, the micro-sweep's chi = 16)
        p = std::min(std::max((int)round_up(2 * k + c->flag[F_SVD_OVS], 16), (int)round_up(k + 32, 16)), PMAX);
        if (p < k + 32 || n <= p || m < p) return MPST_OK;            // too little oversampling: full Jacobi
    }
    if (c->svd_prepare_only && mode != 2) return MPST_OK;              // only the subspace rounds are graphs
    // Gram paths, opt-in variant (MPST_SVD_GRAMREG=1): the Gram matrix goes through the same Cholesky + register-resident
    // Jacobi as the Rayleigh-Ritz step of the subspace path (order rounded up to a multiple of 16, zero padded) instead of
    // the shared-memory Jacobi on the columns of H.  5x faster at order 96 and identical on full-rank bonds, but NOT the
    // default: on a numerically rank-deficient H (first-sweep bonds of a small training set: 20-28 of 80 pivots at
    // rounding level) deflating a pivot d_k drops that column's off-diagonal entries, which are sqrt(d_k h_jj)-sized
    // (an un-pivoted Cholesky of a semi-definite matrix), and the Gram paths have no residual check to catch it
    // (teacher-forced A/B on the 67-sample reference case: truncated products 5e-4 apart, test accuracy after one
    // sweep 0.918 against 0.972 / 0.973 for the oracle / the default).
    const bool gram_reg = mode != 2 && c->flag[F_SVD_GRAMREG];
    if (gram_reg) p = std::max(32, (int)round_up(mode == 0 ? n : m, 16));
    const size_t need = (size_t)2 * n * p + (size_t)2 * m * p + (size_t)(MAXSPLIT + 4) * p * p + (size_t)2 * n * k +
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
    double* Lm = Wk + (size_t)p * p;                                   // Cholesky factor of the Ritz matrix
    double* T2 = Lm + (size_t)p * p;                                   // n x k
    double* Uk = T2 + (size_t)n * k;                                   // m x k
    double* Vs = Uk + (size_t)m * k;                                   // n x k staging block of V_k (subspace mode)
    double* ev = Vs + (size_t)n * k;                                   // p, then Psorted p
    int* status = c->iscal + 8;
    unsigned long long* resbits = reinterpret_cast<unsigned long long*>(c->iscal + 10);
    const size_t eig_smem = 2 * sizeof(double) * (size_t)p * (p + 1) + sizeof(double) * p + sizeof(int) * p + 16;
    if (!c->svd_prepare_only) CUDA_TRY(c, cudaMemsetAsync(status, 0, sizeof(int), c->stream));
    const bool dbg = c->flag[F_SVD_DEBUG] != 0;
    if (dbg) cudaStreamSynchronize(c->stream);
    const auto t0 = std::chrono::steady_clock::now();

    auto finish = [&](const char* what, int iters, double res) -> int {
        *chi_new = c->hiscal[0];
        c->last[L_SVD_PATH] = mode == 0 ? 1 : (mode == 1 ? 2 : 3);
        c->last[L_SVD_ITERS] = iters;
        c->last[L_SVD_ITERS_SUM] += iters;
        if (c->st_on) {
            // one pass of a deflated two-pass split: hand the triplets to the staging blocks, the caller assembles the cores
            const size_t col0 = (size_t)c->st_col0;
            if (*chi_new > 0) {
                CUDA_TRY(c, cudaMemcpyAsync(c->stU + (size_t)m * col0, Uk, sizeof(double) * (size_t)m * (*chi_new), cudaMemcpyDeviceToDevice, c->stream));
                CUDA_TRY(c, cudaMemcpyAsync(c->stV + (size_t)n * col0, Vs, sizeof(double) * (size_t)n * (*chi_new), cudaMemcpyDeviceToDevice, c->stream));
                CUDA_TRY(c, cudaMemcpyAsync(c->stP + col0, ev + p, sizeof(double) * (*chi_new), cudaMemcpyDeviceToDevice, c->stream));
            }
            if (dbg) fprintf(stderr, "[svd %s, pass at column %d] m=%d n=%d p=%d iters=%d chi=%d residual=%.2e\n", what, c->st_col0, m, n, p, iters, *chi_new, res);
            *done = true;
            return MPST_OK;
        }
        c->last[L_SVD_FAST]++;
        if (mode == 2) CUDA_TRY(c, cudaMemcpyAsync(ortho_core, Vs, sizeof(double) * (size_t)n * (*chi_new), cudaMemcpyDeviceToDevice, c->stream));
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
        if (gram_reg) {
            CUDA_TRY(c, cudaMemsetAsync(Gm, 0, sizeof(double) * (size_t)p * p, c->stream));
            if (mode == 0) TRY(launch_dgemm(c, 1, 0, q, q, m, M, ldm, M, ldm, Gm, p));                      // M^T M, ld = p
            else TRY(launch_dgemm(c, 0, 1, q, q, n, M, ldm, M, ldm, Gm, p));                               // M M^T
            launch_chol_inv(p, Gm, 1, Ri, Lm, status, c->stream, !c->flag[F_SVD_CHOLSEQ]);                 // H = L L^T
            c->launches++;
            if (!launch_sym_eig_reg(p, Lm, Wm, ev, status, c->stream))
                if (launch_sym_eig<true>(p, eig_smem, Lm, 1, p, Wm, ev, status, c->stream)) CUDA_TRY(c, cudaGetLastError());
        } else {
        if (mode == 0) TRY(launch_dgemm_splitk(c, 1, 0, q, q, m, M, ldm, M, ldm, Gm, MAXSPLIT, &splits));   // M^T M
        else TRY(launch_dgemm_splitk(c, 0, 1, q, q, n, M, ldm, M, ldm, Gm, MAXSPLIT, &splits));            // M M^T
        if (launch_sym_eig<false>(p, eig_smem, Gm, splits, q, Wm, ev, status, c->stream)) CUDA_TRY(c, cudaGetLastError());
        }
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
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal, c->iscal, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));   // chi_new, non-finite flag
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
        // Gram matrix through the auto split-K GEMM (partials summed by a wide reduce kernel): the single-CTA
        // Cholesky then reads one p x p matrix instead of ~15 partials (its load phase was as long as its pivot loop)
        TRY(launch_dgemm(c, 1, 0, p, p, rows, X, rows, X, rows, Gm, p));
        launch_chol_inv(p, Gm, 1, Ri, nullptr, status, c->stream, !c->flag[F_SVD_CHOLSEQ]);
        c->launches++;
        TRY(launch_dgemm(c, 0, 0, rows, p, p, X, rows, Ri, p, out, rows));
        return MPST_OK;
    };
    // the random start needs no orthonormalisation: an n x p matrix of iid entries is well conditioned
    // (kappa ~ (sqrt(n)+sqrt(p))/(sqrt(n)-sqrt(p))) and only its span matters
    const int max_rounds = 3;
    int iters_done = 0;
    // Column-scaling shortcut (training sweeps only): from the second iteration on Z = M Q is only column-normalised
    // instead of orthonormalised.  If that ever makes a Cholesky pivot break down on a bond, the bond is flagged and
    // this call restarts with full orthonormalisation, so the shortcut can cost time but never correctness.
    const int slot0 = (c->svd_slot >= 0 && c->svd_slot < (int)c->svd_its.size()) ? c->svd_slot : -1;
    // a breakdown (rank-deficient bond early in training) sends the bond's next three visits down the serial loop with full
    // orthonormalisation, then the shortcuts are tried again
    if (slot0 >= 0 && c->svd_nohalf[slot0] > 0 && !c->svd_prepare_only) c->svd_nohalf[slot0]--;
    const bool penalised = slot0 >= 0 && c->svd_nohalf[slot0] > (c->svd_prepare_only ? 1 : 0);
    bool half_orth = slot0 >= 0 && !penalised && !c->flag[F_SVD_NOHALF];
    const int half_from = c->flag[F_SVD_HALF_FROM];
    // overlapped loop (default, also for the stand-alone mpst_bond_split): any Cholesky breakdown restarts the call in
    // the serial form with full orthonormalisation of both blocks, and a training bond remembers it
    bool overlap = !c->flag[F_SVD_SERIAL] && !c->flag[F_SVD_NOHALF] && c->stream2 != nullptr && !penalised;

    // Everything one round puts on the stream(s), from the start block to the read-back of {chi_new, flags, status,
    // residual}.  It touches only workspace buffers (V_k goes to the staging block `Vs`), so for a given shape and
    // iteration count it is the same sequence of ~100 launches on every bond: it is captured once into a CUDA graph and
    // replayed (one launch call instead of ~100; in a sweep the GPU is idle when the split starts and the host could
    // not enqueue the small kernels as fast as they ran).  `fresh`: first round of a call (random start block).
    auto enqueue_round = [&](bool fresh, int niter, int iters_before) -> int {
        double* qa = Qa;
        double* qb = Qb;
        int itd = iters_before;
        if (fresh) {
            CUDA_TRY(c, cudaMemsetAsync(status, 0, sizeof(int), c->stream));
            const int64_t tot = (int64_t)n * p;
            rand_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(qa, n, p, n, 0x5DEECE66Dull + (uint64_t)m * 131 + n);
            c->launches++;
        }
        if (overlap) {
            // Overlapped form of the same iteration.  With Y_i = M^T M Q_{i-1} and Q_i = Y_i R_i^-1 (Cholesky-QR), the next
            // block is Y_{i+1} = M^T M Q_i = (M^T (M Y_i)) R_i^-1: the two big products do not need R_i, so the Gram
            // matrix and its Cholesky inverse (one SM, latency bound, as long as both products together) run on the
            // side stream next to them and R_i^-1 is applied afterwards.  Q_i itself is only formed after the last
            // iteration.  The left block Z is never re-orthonormalised (its columns are graded but well separated,
            // and Cholesky does not care about column scales).
            int todo = niter;
            if (fresh) {
                TRY(launch_dgemm(c, 0, 0, m, p, n, M, ldm, qa, n, Za, m));        // Z = M Q0 (random start)
                TRY(cholqr(Za, Zb, m));
                TRY(launch_dgemm(c, 1, 0, n, p, m, M, ldm, Zb, m, qb, n));        // Y_1 = M^T orth(Z)
            } else {
                TRY(launch_dgemm(c, 1, 0, n, p, m, M, ldm, Za, m, qb, n));        // next round: Y = M^T (M Q) of the failed Rayleigh-Ritz
            }
            todo--;
            for (; todo > 0; todo--) {
                CUDA_TRY(c, cudaEventRecord(c->ev_fork, c->stream));
                CUDA_TRY(c, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
                TRY(launch_dgemm_on(c, 1, 1, 0, p, p, n, qb, n, qb, n, Gm, p));   // side: G = Y^T Y, R^-1
                launch_chol_inv(p, Gm, 1, Ri, nullptr, status, c->stream2, !c->flag[F_SVD_CHOLSEQ]);
                c->launches++;
                CUDA_TRY(c, cudaEventRecord(c->ev_join, c->stream2));
                TRY(launch_dgemm(c, 0, 0, m, p, n, M, ldm, qb, n, Za, m));        // main: M Y, M^T (M Y)
                TRY(launch_dgemm(c, 1, 0, n, p, m, M, ldm, Za, m, qa, n));
                CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
                TRY(launch_dgemm(c, 0, 0, n, p, p, qa, n, Ri, p, qb, n));         // Y_{i+1} = (M^T M Y_i) R_i^-1
            }
            TRY(cholqr(qb, qa, n));                                               // Q = orth(Y)
        } else
        for (int it = 0; it < niter; it++, itd++) {
            TRY(launch_dgemm(c, 0, 0, m, p, n, M, ldm, qa, n, Za, m));            // Z = M Q
            if (half_orth && itd >= half_from) {
                colscale_kernel<<<p, 256, 0, c->stream>>>(Za, Zb, m);
                c->launches++;
            } else {
                TRY(cholqr(Za, Zb, m));
            }
            TRY(launch_dgemm(c, 1, 0, n, p, m, M, ldm, Zb, m, qb, n));            // Y = M^T Z
            TRY(cholqr(qb, qa, n));
        }
        // second pass: orthonormal to rounding (the first pass leaves ||Q^T Q - I|| ~ eps kappa(Y)^2), and Z = M Q.
        // The result always lands in Qa (the next round continues from it).
        if (overlap) {
            // Q1 is orthonormal to ~1e-9 already, so Z = M Q = (M Q1) R2^-1 loses nothing: the big product runs next to
            // the Gram / Cholesky chain of the second pass
            CUDA_TRY(c, cudaEventRecord(c->ev_fork, c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
            TRY(launch_dgemm_on(c, 1, 1, 0, p, p, n, qa, n, qa, n, Gm, p));
            launch_chol_inv(p, Gm, 1, Ri, nullptr, status, c->stream2, !c->flag[F_SVD_CHOLSEQ]);
            c->launches++;
            CUDA_TRY(c, cudaEventRecord(c->ev_join, c->stream2));
            TRY(launch_dgemm(c, 0, 0, m, p, n, M, ldm, qa, n, Zb, m));            // M Q1
            CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
            TRY(launch_dgemm(c, 0, 0, m, p, p, Zb, m, Ri, p, Za, m));             // Z = (M Q1) R2^-1
            TRY(launch_dgemm(c, 0, 0, n, p, p, qa, n, Ri, p, qb, n));             // Q = Q1 R2^-1
            CUDA_TRY(c, cudaMemcpyAsync(qa, qb, sizeof(double) * (size_t)n * p, cudaMemcpyDeviceToDevice, c->stream));
        } else {
            TRY(cholqr(qa, qb, n));
            CUDA_TRY(c, cudaMemcpyAsync(qa, qb, sizeof(double) * (size_t)n * p, cudaMemcpyDeviceToDevice, c->stream));
            TRY(launch_dgemm(c, 0, 0, m, p, n, M, ldm, qa, n, Za, m));            // Z = M Q
        }
        // Rayleigh-Ritz
        TRY(launch_dgemm(c, 1, 0, p, p, m, Za, m, Za, m, Gm, p));                 // H = Z^T Z
        // H = L L^T (Cholesky in registers, breakdown -> status -> exact fallback), then Jacobi on the columns of L
        launch_chol_inv(p, Gm, 1, Ri, Lm, status, c->stream, !c->flag[F_SVD_CHOLSEQ]);
        // register-resident Jacobi (sym_eig_reg.cu) for p = 16k; the shared-memory solver otherwise / on request
        if (c->flag[F_SVD_EIGSMEM] || !launch_sym_eig_reg(p, Lm, Wm, ev, status, c->stream))
            if (launch_sym_eig<true>(p, eig_smem, Lm, 1, p, Wm, ev, status, c->stream)) CUDA_TRY(c, cudaGetLastError());
        c->launches++;
        ritz_trunc_kernel<<<1, 256, 0, c->stream>>>(ev, p, n, trace_dev, chi_max, cutoff, c->perm, ev + p, c->iscal, c->st_on ? c->st_kept : nullptr);
        gather_cols_kernel<<<(p * k + 255) / 256, 256, 0, c->stream>>>(Wm, p, c->perm, k, Wk);
        c->launches += 3;
        TRY(launch_dgemm(c, 0, 0, n, k, p, qa, n, Wk, p, Vs, n));                 // V_k = Q W_k
        TRY(launch_dgemm(c, 0, 0, m, k, p, Za, m, Wk, p, Uk, m));                 // U_k S_k = Z W_k
        // residual of the kept Ritz pairs: M^T (M v_i) - sigma_i^2 v_i
        TRY(launch_dgemm(c, 1, 0, n, k, m, M, ldm, Uk, m, T2, n));
        CUDA_TRY(c, cudaMemsetAsync(resbits, 0, sizeof(unsigned long long), c->stream));
        residual_kernel<<<k, 256, 0, c->stream>>>(T2, Vs, ev + p, c->iscal, n, resbits, (c->st_on && c->st_kept) ? c->stP : nullptr);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
        // one read-back: chi_new, non-finite flag [0..1], Cholesky / Jacobi status [8..9], residual bits [10..11]
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal, c->iscal, 12 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        return MPST_OK;
    };
    // graph cache: key = everything the captured sequence depends on (shapes, counts, switches, buffer addresses)
    auto run_round = [&](bool fresh, int niter, int iters_before) -> int {
        const bool graphable = !dbg && !c->flag[F_SVD_NOGRAPH] && overlap;
        if (!graphable) return enqueue_round(fresh, niter, iters_before);
        // split-K workspaces must not grow inside a capture
        TRY(ensure_buf(c, &c->gws, &c->gwscap, std::max((size_t)16 * std::max(m, n) * p, (size_t)c->sm_count * p * p)));
        TRY(ensure_buf(c, &c->gws2, &c->gws2cap, (size_t)c->sm_count * p * p));
        uint64_t cbits;
        memcpy(&cbits, &cutoff, sizeof cbits);
        std::vector<int64_t> key = {m, n, (int64_t)ldm, p, k, niter, fresh ? 1 : 0, chi_max, (int64_t)cbits, C,
                                    (int64_t)(uintptr_t)M, (int64_t)(uintptr_t)trace_dev, (int64_t)(uintptr_t)c->sub,
                                    (int64_t)(uintptr_t)c->gws, (int64_t)(uintptr_t)c->gws2, c->flag[F_SVD_CHOLSEQ], c->flag[F_SVD_EIGSMEM], c->st_on ? 1 : 0,
                                    (int64_t)(uintptr_t)(c->st_on ? c->st_kept : nullptr)};
        SvdGraph* g = nullptr;
        for (auto& e : c->svd_graphs) if (e.key == key) { g = &e; break; }
        if (!g) {
            const int64_t l0 = c->launches;
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                c->flag[F_SVD_NOGRAPH] = 1;
                return enqueue_round(fresh, niter, iters_before);
            }
            const int rc = enqueue_round(fresh, niter, iters_before);
            const cudaError_t ee = cudaStreamEndCapture(c->stream, &graph);
            const int64_t nl = c->launches - l0;
            c->launches = l0;
            cudaGraphExec_t exec = nullptr;
            if (rc != MPST_OK || ee != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
                cudaGetLastError();
                if (graph) cudaGraphDestroy(graph);
                c->flag[F_SVD_NOGRAPH] = 1;                                        // capture not possible here: plain launches from now on
                c->err.clear();
                return enqueue_round(fresh, niter, iters_before);
            }
            cudaGraphDestroy(graph);
            if (c->svd_graphs.size() >= 24) {                                      // drop the least recently used
                size_t lru = 0;
                for (size_t i = 1; i < c->svd_graphs.size(); i++) if (c->svd_graphs[i].last_use < c->svd_graphs[lru].last_use) lru = i;
                cudaGraphExecDestroy((cudaGraphExec_t)c->svd_graphs[lru].exec);
                c->svd_graphs.erase(c->svd_graphs.begin() + lru);
            }
            c->svd_graphs.push_back({key, (void*)exec, nl, 0});
            g = &c->svd_graphs.back();
        }
        g->last_use = ++c->seg_clock;
        if (c->svd_prepare_only) {
            // called ahead of the split (while the gradient kernel runs): push the graph's launch state to the device on
            // the idle side stream, so that the launch itself finds nothing left to do on the host
            CUDA_TRY(c, cudaGraphUpload((cudaGraphExec_t)g->exec, c->stream2));
            CUDA_TRY(c, cudaEventRecord(c->ev_upload, c->stream2));
            c->svd_uploaded = g->exec;
            return MPST_OK;
        }
        if (c->svd_uploaded == g->exec) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_upload, 0));
        c->svd_uploaded = nullptr;
        CUDA_TRY(c, cudaGraphLaunch((cudaGraphExec_t)g->exec, c->stream));
        c->launches += g->launches;
        return MPST_OK;
    };
    // Rounds 0..2 are the measured policy of the training sweeps (first, then 3 + 3 iterations).  Where that policy used
    // to hand the bond to the exact Jacobi (a slowly decaying spectrum on a big matrix: 3.3 s at 6144 x 3072 against
    // 0.3 ms per iteration), the observed contraction rate now decides: if it predicts convergence within `adaptive_budget`
    // more iterations the loop goes on with plain launches (no graph: these iteration counts do not repeat).
    double prev_res = -1.0;
    int next_n = 3, adaptive_budget = 48;
    bool adaptive = false;
restart:
    prev_res = -1.0; next_n = 3; adaptive = false;
    for (int round = 0; round < max_rounds + 8; round++) {
        // iterations of the first round: 7 (5 when the subspace has 2k columns) unless this bond's previous visits
        // showed what reaches the residual bound (trained spectra change slowly from sweep to sweep)
        const int slot = (c->svd_slot >= 0 && c->svd_slot < (int)c->svd_its.size()) ? c->svd_slot : -1;
        int first = c->flag[F_SVD_IT] > 0 ? c->flag[F_SVD_IT] : (p >= 2 * k ? 5 : c->flag[F_SVD_FIRST]);
        const bool hinted = slot >= 0 && c->svd_its[slot] == 0 && c->flag[F_SVD_IT] <= 0 && c->svd_hint_its > 0 &&
                            c->svd_hint_m == m && c->svd_hint_n == n;
        if (slot >= 0 && c->svd_its[slot] > 0 && c->flag[F_SVD_IT] <= 0) first = c->svd_its[slot];
        else if (hinted) first = std::max(c->svd_hint_its, c->svd_floor[slot]);
        const int niter = round == 0 ? first : next_n;
        if (c->svd_prepare_only) {
            if (!dbg && !c->flag[F_SVD_NOGRAPH] && overlap) TRY(run_round(true, niter, 0));
            return MPST_OK;
        }
        if (adaptive) TRY(enqueue_round(false, niter, iters_done));
        else TRY(run_round(iters_done == 0, niter, iters_done));
        iters_done += niter;
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        double res;
        memcpy(&res, c->hiscal + 10, sizeof(double));
        if (c->hiscal[8] != 0 || !(res == res)) {
            if (half_orth || overlap) {                                            // retry once without the shortcuts
                if (dbg) fprintf(stderr, "[svd subspace] m=%d n=%d breakdown with column scaling -> restart with full orth\n", m, n);
                if (slot0 >= 0) {
                    // back-off: 1, 2, 4, 8 visits down the serial loop for a bond that keeps breaking down (a genuinely
                    // rank-deficient bond); a rank-deficient first sweep (chi_init start) costs the next visit only
                    if ((int)c->svd_pen.size() != (int)c->svd_nohalf.size()) c->svd_pen.assign(c->svd_nohalf.size(), 0);
                    c->svd_pen[slot0] = (char)std::min(8, std::max(1, 2 * (int)c->svd_pen[slot0]));
                    c->svd_nohalf[slot0] = (char)(c->svd_pen[slot0] + 1);
                }
                c->last[L_SVD_RESTARTS]++;
                half_orth = false;
                overlap = false;
                iters_done = 0;
                goto restart;
            }
            if (dbg) fprintf(stderr, "[svd subspace] m=%d n=%d breakdown status=%d -> full Jacobi\n", m, n, c->hiscal[8]);
            return MPST_OK;
        }
        if (slot >= 0 && c->flag[F_SVD_PROBE]) {
            // round-1 policy (kept for comparison): try one iteration fewer whenever the residual passed with a little margin
            if (round == 0 && res <= 5e-14) {
                if (res <= 1.5e-14 && first - 1 >= std::max(3, c->svd_floor[slot])) c->svd_its[slot] = first - 1;
                else c->svd_its[slot] = first;
                c->svd_hint_m = m; c->svd_hint_n = n;
                c->svd_hint_its = (res <= 1.5e-14 && first - 1 >= std::max(3, c->svd_hint_floor)) ? first - 1 : first;
            } else if (round == 0) {
                c->svd_floor[slot] = first + 1;
                c->svd_its[slot] = first + 1;
                if (c->svd_hint_m == m && c->svd_hint_n == n) {
                    c->svd_hint_floor = std::max(c->svd_hint_floor, first + 1);
                    c->svd_hint_its = std::max(c->svd_hint_its, first + 1);
                }
            }
        } else if (slot >= 0) {
            // An extra iteration costs two products (~80 us at 2048 x 1024), a failed round a whole second Rayleigh-Ritz
            // (~1 ms): the count is kept where the residual sits at the rounding floor, raised as soon as the margin
            // to the bound gets thin, and lowered only after the bond has sat at the floor for three visits in a row
            // (a level that failed on this bond is never tried again).
            if ((int)c->svd_calm.size() != (int)c->svd_its.size()) c->svd_calm.assign(c->svd_its.size(), 0);
            if (round == 0 && res <= 5e-14) {
                int next = first;
                if (res > 2e-14) { next = std::min(first + 1, 12); c->svd_calm[slot] = 0; }
                else if (res <= 1.5e-15) {
                    if (++c->svd_calm[slot] >= 3 && first - 1 >= std::max(3, c->svd_floor[slot])) { next = first - 1; c->svd_calm[slot] = 0; }
                } else c->svd_calm[slot] = 0;
                c->svd_its[slot] = next;
                c->svd_hint_m = m; c->svd_hint_n = n;
                c->svd_hint_its = std::max(first, next);
            } else if (round == 0) {
                c->svd_calm[slot] = 0;
                c->svd_floor[slot] = first + 1;
                c->svd_its[slot] = first + 1;
                if (c->svd_hint_m == m && c->svd_hint_n == n) c->svd_hint_its = std::max(c->svd_hint_its, first + 1);
            }
        }
        if (res <= 5e-14) {
            if (overlap && slot0 >= 0 && (int)c->svd_pen.size() == (int)c->svd_nohalf.size()) c->svd_pen[slot0] = 0;
            if (!overlap) c->last[L_SVD_SERIAL]++;
            return finish("subspace", iters_done, res);
        }
        if (round == 0) c->last[L_SVD_ROUND2]++;
        if (round < max_rounds - 1 && res <= (round == 0 ? 1e-5 : 1e-10)) { prev_res = res; next_n = 3; continue; }
        // contraction per iteration: from the last two residuals, or (round 0) from the start at ~1
        const double rate = (round == 0 || !(prev_res > 0.0)) ? pow(res, 1.0 / niter) : pow(res / prev_res, 1.0 / niter);
        const int need = (rate > 0.0 && rate < 0.85) ? (int)ceil(log(2e-14 / res) / log(rate)) : 1 << 20;
        if ((round == 0 && res > 1e-3) || need > adaptive_budget) {               // spectrum too flat: full Jacobi
            if (dbg) fprintf(stderr, "[svd subspace] m=%d n=%d residual %.2e after %d its (rate %.2f) -> full Jacobi\n", m, n, res, iters_done, rate);
            return MPST_OK;
        }
        next_n = std::min(std::max(need, 3), 16);
        adaptive_budget -= next_n;
        adaptive = true;
        prev_res = res;
        if (dbg) fprintf(stderr, "[svd subspace] m=%d n=%d residual %.2e after %d its, rate %.2f per iteration -> %d more\n", m, n, res, iters_done, rate, next_n);
    }
    return MPST_OK;                                                                // not converged: full Jacobi
}

// Entry point.  chi_max <= 80 (or a Gram-path shape): one call of subspace_core.  Larger chi_max: the single-CTA Cholesky
// and Rayleigh-Ritz kernels stop at 128 columns and a subspace needs chi_max + 32, so the split runs as TWO passes of
// the same machinery (each on <= 112 columns, the shape the north-star bonds use):
//   pass 1 : the 64 leading triplets of M (truncation rule with the global trace; fewer than 64 kept => finished),
//   deflate: M2 = M - (U1 S1) V1^T,
//   pass 2 : up to chi_max - 64 triplets of M2; the truncation rule continues where pass 1 stopped (weight already kept
//            = sum sigma_i^2 of pass 1, this pass may keep nothing); residuals measured against sigma_1^2 of M.
// M2 carries pass 1's rounding (eps sigma_1 per entry): sigma and the truncated product agree with LAPACK to ~1e-15
// sigma_1, V2 is orthogonal to V1 to eps sigma_1 / sigma_j (1e-12 measured at the cutoff edge).  Any pass that does not
// converge leaves *done = false and the caller runs the exact Jacobi on the untouched M.
int svd_subspace_device(mpst_ctx* c, const double* M, int64_t ldm, int m, int n, int C, int chi_max, double cutoff,
                        const double* trace_dev, double* label_core, double* ortho_core, int* chi_new,
                        double* sigma_host, bool* done) {
    *done = false;
    const int k = std::min(chi_max, std::min(m, n));
    const int K1 = 64, PMAX = 112;
    const bool two_pass = k > 80 && n > PMAX && m > PMAX && !c->flag[F_SVD_NO2PASS] && !c->flag[F_SVD_NOSUB] && cutoff >= 1e-12;
    if (!two_pass) return subspace_core(c, M, ldm, m, n, C, chi_max, cutoff, trace_dev, label_core, ortho_core, chi_new, sigma_host, done);
    if (c->svd_prepare_only) return MPST_OK;
    const size_t szU = (size_t)round_up((int64_t)m * k, 32), szV = (size_t)round_up((int64_t)n * k, 32), szP = (size_t)round_up(k + 8, 32);
    TRY(ensure_buf(c, &c->stbuf, &c->stbufcap, szU + szV + szP + (size_t)m * n));
    c->stU = c->stbuf;
    c->stV = c->stU + szU;
    c->stP = c->stV + szV;
    double* kept = c->stP + k;
    double* M2 = c->stP + szP;
    const int slot_saved = c->svd_slot;                    // the per-bond iteration history belongs to single-pass splits
    c->svd_slot = -1;
    c->st_on = true;
    c->st_col0 = 0;
    c->st_kept = nullptr;
    int chi1 = 0, chi2 = 0;
    bool d1 = false, d2 = true;
    int rc = subspace_core(c, M, ldm, m, n, C, K1, cutoff, trace_dev, nullptr, nullptr, &chi1, nullptr, &d1);
    if (rc == MPST_OK && d1 && chi1 == K1) {
        d2 = false;
        sum_first_kernel<<<1, 32, 0, c->stream>>>(c->stP, K1, kept);
        c->launches++;
        rc = launch_dgemm(c, 0, 1, m, n, K1, c->stU, m, c->stV, n, M2, m);                 // (U1 S1) V1^T
        if (rc == MPST_OK) {
            deflate_kernel<<<2 * c->sm_count, 256, 0, c->stream>>>(M, ldm, M2, m, n);
            c->launches++;
            c->st_col0 = K1;
            c->st_kept = kept;
            rc = subspace_core(c, M2, m, m, n, C, k - K1, cutoff, trace_dev, nullptr, nullptr, &chi2, nullptr, &d2);
        }
    }
    c->st_on = false;
    c->st_kept = nullptr;
    c->svd_slot = slot_saved;
    if (rc != MPST_OK) return rc;
    if (!d1 || !d2) return MPST_OK;                        // not converged: exact Jacobi on M
    const int chi = chi1 + chi2;
    *chi_new = chi;
    c->hiscal[0] = chi;
    CUDA_TRY(c, cudaMemcpyAsync(ortho_core, c->stV, sizeof(double) * (size_t)n * chi, cudaMemcpyDeviceToDevice, c->stream));
    scatter_label_n_kernel<<<2 * c->sm_count, 256, 0, c->stream>>>(c->stU, m, m / C, chi, c->iscal, label_core);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    if (sigma_host) {
        std::vector<double> tmp(chi);
        CUDA_TRY(c, cudaMemcpyAsync(tmp.data(), c->stP, sizeof(double) * chi, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        for (int q = 0; q < chi; q++) sigma_host[q] = sqrt(tmp[q]);
    }
    c->last[L_SVD_PATH] = 3;
    c->last[L_SVD_FAST]++;
    if (chi2 > 0 || chi1 == K1) c->last[L_SVD_TWOPASS]++;
    *done = true;
    return MPST_OK;
}
