// K5: on-device truncated SVD of the updated bond tensor (replaces decomposeBT ->
// ITensors.svd (LAPACK gesdd) + NDTensors truncate!, reference
// Training/RealRealHighDimension.jl:146-203).
//
// Algorithm: one-sided block Jacobi (Hestenes) on the stacked matrix S = [M ; I_n]
// (M is m x n, rows carry the class/label index, columns are the (site, link) pair that stays
// orthonormal).  Column blocks of 16 are paired by a round-robin tournament; for every pair
//   (1) jac_gram   : 32x32 Gram matrix of the pair's columns (M rows only), split over row slices
//   (2) jac_solve  : cyclic two-sided Jacobi on the 32x32 Gram in shared memory -> rotation W
//   (3) jac_apply  : S[:, pair] <- S[:, pair] * W   (M rows and the accumulated V rows)
// At convergence the M rows hold U*diag(sigma) (exactly the "moving" core the reference builds as
// U*S) and the I rows hold V (the orthonormal core).  No Gram matrix of the full problem is ever
// formed, so small singular values keep full relative accuracy.  Everything is deterministic
// (fixed pairing and summation order), so replicated ranks stay bit-identical.
// Truncation: NDTensors rule on P = sigma^2 (maxdim, then relative cutoff on the *sum* of
// discarded weight), evaluated on the device.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "mpst_common.cuh"

int svd_subspace_device(mpst_ctx* c, const double* M, int64_t ldm, int m, int n, int C, int chi_max, double cutoff,
                        const double* trace_dev, double* label_core, double* ortho_core, int* chi_new,
                        double* sigma_host, bool* done);

namespace {
constexpr int GR = 64;                   // rows per Gram chunk

__device__ __forceinline__ void rr_pair(int nb, int st, int k, int& I, int& J) {
    int a, b;
    if (k == 0) { a = nb - 1; b = st; }
    else { a = (st + k) % (nb - 1); b = (st - k + (nb - 1)) % (nb - 1); }
    I = min(a, b);
    J = max(a, b);
}

// going_left : col = q (n = Dr), rows (c, p): S = B[c][p + Dl*q]
// going right: col = p (n = Dl), rows (c, q): S = B[c][p + Dl*q]
__global__ void __launch_bounds__(256)
jac_load_kernel(const double* __restrict__ B, double* __restrict__ S, int going_left, int Dl, int Dr,
                int C, int m, int n, int npad, int64_t ld, const double* __restrict__ norm2, int use_scale) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t mm = (int64_t)m + n;
    if (e >= mm * npad) return;
    const int col = (int)(e / mm);
    const int r = (int)(e - (int64_t)col * mm);
    double v = 0.0;
    if (col < n) {
        if (r < m) {
            const int Dx = going_left ? Dl : Dr;
            const int c = r / Dx, xx = r - c * Dx;
            const int p = going_left ? xx : col, q = going_left ? col : xx;
            v = B[(size_t)c * Dl * Dr + (size_t)p + (size_t)Dl * q];
            if (use_scale) v *= 1.0 / sqrt(*norm2);
        } else {
            v = (r - m == col) ? 1.0 : 0.0;
        }
    }
    S[(size_t)col * ld + r] = v;
}

// Gram matrix of one column-block pair over a row slice.  256 threads, each a (PB/16)x(PB/16) block.
template <int PB>
__global__ void __launch_bounds__(256)
jac_gram_kernel(const double* __restrict__ S, int64_t ld, int m, int nb, int st, int rsplit,
                double* __restrict__ gpart) {
    constexpr int JB = PB / 2, R = PB / 16;
    __shared__ double T[PB][GR + 1];
    int I, J;
    rr_pair(nb, st, blockIdx.x, I, J);
    const int split = blockIdx.y;
    const int nchunks = (m + GR - 1) / GR;
    const int c0 = (int)(((int64_t)nchunks * split) / rsplit), c1 = (int)(((int64_t)nchunks * (split + 1)) / rsplit);
    const int tid = threadIdx.x;
    const int u = tid & 15, v = tid >> 4;
    double acc[R][R];
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
        for (int j = 0; j < R; j++) acc[i][j] = 0.0;
    const int lr = tid & 63, lg = tid >> 6;                 // loader: row, column group
    for (int ch = c0; ch < c1; ch++) {
        const int r = ch * GR + lr;
#pragma unroll
        for (int k = 0; k < PB / 4; k++) {
            const int cc = lg * (PB / 4) + k;
            const int col = (cc < JB) ? I * JB + cc : J * JB + (cc - JB);
            T[cc][lr] = (r < m) ? S[(size_t)col * ld + r] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < GR; rr++) {
            double x[R], y[R];
#pragma unroll
            for (int i = 0; i < R; i++) { x[i] = T[R * u + i][rr]; y[i] = T[R * v + i][rr]; }
#pragma unroll
            for (int i = 0; i < R; i++)
#pragma unroll
                for (int j = 0; j < R; j++) acc[i][j] += x[i] * y[j];
        }
        __syncthreads();
    }
    double* g = gpart + ((size_t)blockIdx.x * rsplit + split) * PB * PB;
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
        for (int j = 0; j < R; j++) g[(R * u + i) * PB + R * v + j] = acc[i][j];
}

// Rotation matrix of the pair: `inner_sweeps` cyclic two-sided Jacobi sweeps on the PBxPB Gram matrix in
// shared memory (PB/2 disjoint rotations per round, 16 threads per rotation).  A pair (p,q) is rotated only
// if |a_pq| > tol*sqrt(a_pp a_qq) AND |a_pq| > floor_abs: entries below the absolute floor (a fraction of
// eps*||M||_F^2, far below the truncation cutoff) are rounding-level couplings between numerically null
// directions -- LAPACK's gesdd does not resolve them either -- and chasing them stalls convergence.
// W[v][u] = component v of new column u.
template <int PB>
__global__ void __launch_bounds__(PB * 8)
jac_solve_kernel(const double* __restrict__ gpart, int rsplit, double* __restrict__ wbuf, double tol,
                 double abs_tol, const double* __restrict__ trace_dev, int inner_sweeps,
                 unsigned long long* __restrict__ maxoff_bits) {
    constexpr int NT = PB * 8;
    extern __shared__ double jsm[];
    double (*A)[PB + 1] = reinterpret_cast<double (*)[PB + 1]>(jsm);
    double (*V)[PB + 1] = reinterpret_cast<double (*)[PB + 1]>(jsm + PB * (PB + 1));
    __shared__ int any_rot;
    const int tid = threadIdx.x;
    const double floor_abs = abs_tol * (trace_dev ? *trace_dev : 1.0);
    const double* g = gpart + (size_t)blockIdx.x * rsplit * PB * PB;
    for (int e = tid; e < PB * PB; e += NT) {
        double s = 0.0;
        for (int k = 0; k < rsplit; k++) s += g[(size_t)k * PB * PB + e];
        A[e / PB][e % PB] = s;
        V[e / PB][e % PB] = (e / PB == e % PB) ? 1.0 : 0.0;
    }
    __syncthreads();
    double mo = 0.0;                                        // convergence measure before rotating
    for (int e = tid; e < PB * PB; e += NT) {
        const int r = e / PB, cidx = e % PB;
        if (r < cidx) {
            const double apq = fabs(A[r][cidx]);
            const double den = A[r][r] * A[cidx][cidx];
            if (apq > floor_abs && den > 0.0) mo = fmax(mo, apq * rsqrt(den));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mo = fmax(mo, __shfl_xor_sync(0xffffffffu, mo, o));
    if ((tid & 31) == 0 && mo > 0.0) atomicMax(maxoff_bits, (unsigned long long)__double_as_longlong(mo));
    __syncthreads();

    const int k = tid >> 4, l16 = tid & 15;
    for (int sweep = 0; sweep < inner_sweeps; sweep++) {
        if (tid == 0) any_rot = 0;
        __syncthreads();
        for (int rd = 0; rd < PB - 1; rd++) {
            int p, q;
            rr_pair(PB, rd, k, p, q);
            const double app = A[p][p], aqq = A[q][q], apq = A[p][q];
            double c = 1.0, s = 0.0;
            const double aa = fabs(apq);
            if (aa > floor_abs && aa > tol * sqrt(fabs(app * aqq))) {
                // t = sgn(tau)/(|tau| + sqrt(1+tau^2)), tau = (aqq-app)/(2 apq), written division-light
                const double zeta = aqq - app, beta = 2.0 * apq;
                const double t = (zeta >= 0.0 ? beta : -beta) / (fabs(zeta) + sqrt(zeta * zeta + beta * beta));
                c = rsqrt(1.0 + t * t);
                s = t * c;
                if (l16 == 0) any_rot = 1;
            }
            __syncthreads();
#pragma unroll
            for (int h = 0; h < PB / 16; h++) {             // columns p, q of A and V
                const int r = l16 + 16 * h;
                const double x = A[r][p], y = A[r][q];
                A[r][p] = c * x - s * y;
                A[r][q] = s * x + c * y;
                const double vx = V[r][p], vy = V[r][q];
                V[r][p] = c * vx - s * vy;
                V[r][q] = s * vx + c * vy;
            }
            __syncthreads();
#pragma unroll
            for (int h = 0; h < PB / 16; h++) {             // rows p, q of A
                const int cc = l16 + 16 * h;
                const double x = A[p][cc], y = A[q][cc];
                A[p][cc] = c * x - s * y;
                A[q][cc] = s * x + c * y;
            }
            __syncthreads();
        }
        if (!any_rot) break;
        __syncthreads();
    }
    double* w = wbuf + (size_t)blockIdx.x * PB * PB;
    for (int e = tid; e < PB * PB; e += NT) w[e] = V[e / PB][e % PB];
}

template <int PB>
__global__ void __launch_bounds__(128)
jac_apply_kernel(double* __restrict__ S, int64_t ld, int mm, int nb, int st, const double* __restrict__ wbuf) {
    constexpr int JB = PB / 2;
    __shared__ double W[PB][PB];
    int I, J;
    rr_pair(nb, st, blockIdx.x, I, J);
    const double* w = wbuf + (size_t)blockIdx.x * PB * PB;
    for (int e = threadIdx.x; e < PB * PB; e += 128) W[e / PB][e % PB] = w[e];
    __syncthreads();
    const int r = blockIdx.y * 128 + threadIdx.x;
    if (r >= mm) return;
    double x[PB];
#pragma unroll
    for (int v = 0; v < PB; v++) {
        const int col = (v < JB) ? I * JB + v : J * JB + (v - JB);
        x[v] = S[(size_t)col * ld + r];
    }
#pragma unroll 2
    for (int u = 0; u < PB; u++) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int v = 0; v < PB; v += 2) { s0 += x[v] * W[v][u]; s1 += x[v + 1] * W[v + 1][u]; }
        const int col = (u < JB) ? I * JB + u : J * JB + (u - JB);
        S[(size_t)col * ld + r] = s0 + s1;
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused sweep: one cooperative launch runs all nb-1 steps of a Jacobi sweep with fine-grained
// dataflow synchronisation instead of 3 kernel launches per step.  CTA (k, r) owns pair slot k and row
// group r.  Per step: wait until the two column blocks of its pair have been updated (same row group)
// by the previous step -> partial Gram over its rows -> barrier among the R CTAs of the slot -> every
// CTA sums the partials in the same order and runs the 32x32 inner sweep redundantly (bit-identical W,
// no broadcast) -> rotates its rows -> publishes the two blocks' step counters.
// All waits are bounded spins on monotone counters; on timeout an error flag stops every CTA (no hang).
constexpr int FPB = 32, FJB = 16;

struct FusedArgs {
    double* S; int64_t ld; int m, mm, nb, R, rows_per;
    double* gpart;                  // [npairs][R][32*32]
    int* done;                      // [nb][R] steps applied
    int* garr;                      // [npairs] arrivals
    int* errflag;
    unsigned long long* maxoff;
    double tol, abs_tol; const double* trace_dev; int inner; int skip;
    const double* theta;            // stopping test looks only at pairs with max(a_pp, a_qq) >= *theta
};

__device__ __forceinline__ bool spin_until(volatile int* p, int target, int* errflag) {
    unsigned spins = 0;
    while (*p < target) {
        if ((++spins & 0xfff) == 0) {
            if (*(volatile int*)errflag) return false;
            if (spins > (1u << 22)) { atomicExch(errflag, 1); return false; }
        }
    }
    return true;
}

__global__ void __launch_bounds__(256)
jac_sweep_fused_kernel(FusedArgs a) {
    __shared__ double A[FPB][FPB + 1];
    __shared__ double V[FPB][FPB + 1];
    __shared__ double T[FPB][GR + 1];
    __shared__ int any_rot, ok_flag;
    const int tid = threadIdx.x;
    const int k = blockIdx.x / a.R, r = blockIdx.x % a.R;
    const int row0 = r * a.rows_per, row1 = min(a.mm, row0 + a.rows_per);
    const int grow1 = min(row1, a.m);                       // Gram rows: the M part only
    const double floor_abs = a.abs_tol * (a.trace_dev ? *a.trace_dev : 1.0);
    const int u = tid & 15, v = tid >> 4;
    const int lr = tid & 63, lg = tid >> 6;
    const int kk = tid >> 4, l16 = tid & 15;
    for (int st = 0; st < a.nb - 1; st++) {
        int I, J;
        rr_pair(a.nb, st, k, I, J);
        if (tid == 0) {
            bool ok = spin_until(a.done + I * a.R + r, st, a.errflag) && spin_until(a.done + J * a.R + r, st, a.errflag);
            __threadfence();
            ok_flag = ok ? 1 : 0;
        }
        __syncthreads();
        if (!ok_flag) return;
        // ---- partial Gram over rows [row0, grow1) ----
        double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
        for (int base = row0; base < ((a.skip & 1) ? row0 : grow1); base += GR) {
            const int rr0 = base + lr;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int cc = lg * 8 + q;
                const int col = (cc < FJB) ? I * FJB + cc : J * FJB + (cc - FJB);
                T[cc][lr] = (rr0 < grow1) ? __ldcg(a.S + (size_t)col * a.ld + rr0) : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int rr = 0; rr < GR; rr++) {
                const double x0 = T[2 * u][rr], x1 = T[2 * u + 1][rr];
                const double y0 = T[2 * v][rr], y1 = T[2 * v + 1][rr];
                a00 += x0 * y0; a01 += x0 * y1; a10 += x1 * y0; a11 += x1 * y1;
            }
            __syncthreads();
        }
        {
            // partials are double-buffered by step parity: a slow sibling may still be summing step st-1
            double* g = a.gpart + ((size_t)(st & 1) * gridDim.x + (size_t)k * a.R + r) * FPB * FPB;
            __stcg(g + (2 * u) * FPB + 2 * v, a00);
            __stcg(g + (2 * u) * FPB + 2 * v + 1, a01);
            __stcg(g + (2 * u + 1) * FPB + 2 * v, a10);
            __stcg(g + (2 * u + 1) * FPB + 2 * v + 1, a11);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicAdd(a.garr + k, 1);
            bool ok = spin_until(a.garr + k, (st + 1) * a.R, a.errflag);
            __threadfence();
            ok_flag = ok ? 1 : 0;
        }
        __syncthreads();
        if (!ok_flag) return;
        // ---- sum partials (fixed order) ----
        {
            const double* g = a.gpart + ((size_t)(st & 1) * gridDim.x + (size_t)k * a.R) * FPB * FPB;
            for (int e = tid; e < FPB * FPB; e += 256) {
                double s = 0.0;
                for (int q = 0; q < a.R; q++) s += __ldcg(g + (size_t)q * FPB * FPB + e);
                A[e / FPB][e % FPB] = s;
                V[e / FPB][e % FPB] = (e / FPB == e % FPB) ? 1.0 : 0.0;
            }
        }
        __syncthreads();
        if (r == 0) {
            double mo = 0.0;
            const double theta = *a.theta;
            for (int e = tid; e < FPB * FPB; e += 256) {
                const int rr = e / FPB, cidx = e % FPB;
                if (rr < cidx) {
                    const double apq = fabs(A[rr][cidx]);
                    const double den = A[rr][rr] * A[cidx][cidx];
                    if (apq > floor_abs && den > 0.0 && fmax(A[rr][rr], A[cidx][cidx]) >= theta) mo = fmax(mo, apq * rsqrt(den));
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) mo = fmax(mo, __shfl_xor_sync(0xffffffffu, mo, o));
            if ((tid & 31) == 0 && mo > 0.0) atomicMax(a.maxoff, (unsigned long long)__double_as_longlong(mo));
        }
        // ---- inner sweep(s) ----
        // Step 0 pairs every block once, so the full 31-round schedule there orthogonalises the pairs inside
        // each block once per sweep; later steps only need the 16 rounds of cross pairs (p in I, q in J).
        for (int sweep = 0; sweep < ((a.skip & 2) ? 0 : a.inner); sweep++) {
            if (tid == 0) any_rot = 0;
            __syncthreads();
            const bool full_rr = (st == 0) || (a.skip & 8);
            const int nrounds = full_rr ? FPB - 1 : FJB;
            for (int rd = 0; rd < nrounds; rd++) {
                int p, q;
                if (full_rr) rr_pair(FPB, rd, kk, p, q);
                else { p = kk; q = FJB + ((kk + rd) & (FJB - 1)); }
                const double app = A[p][p], aqq = A[q][q], apq = A[p][q];
                double c = 1.0, s = 0.0;
                const double aa = fabs(apq);
                if (aa > floor_abs && aa * aa > a.tol * a.tol * fabs(app * aqq)) {
                    // rotation that zeroes a_pq, from cos/sin of the double angle (two rsqrt, no division):
                    // cos2 = |zeta|/r, sin2 = sgn(zeta) beta/r, c = sqrt((1+cos2)/2), s = sin2/(2c)
                    const double zeta = aqq - app, beta = 2.0 * apq;
                    const double inv_r = rsqrt(zeta * zeta + beta * beta);
                    const double cos2 = fabs(zeta) * inv_r;
                    const double sin2 = (zeta >= 0.0 ? beta : -beta) * inv_r;
                    const double c2 = 0.5 + 0.5 * cos2;
                    const double inv_c = rsqrt(c2);
                    c = c2 * inv_c;
                    s = 0.5 * sin2 * inv_c;
                    if (l16 == 0) any_rot = 1;
                }
                __syncwarp();                                // the 16 lanes of a rotation share one warp
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int rr = l16 + 16 * h;
                    const double x = A[rr][p], y = A[rr][q];
                    A[rr][p] = c * x - s * y;
                    A[rr][q] = s * x + c * y;
                    const double vx = V[rr][p], vy = V[rr][q];
                    V[rr][p] = c * vx - s * vy;
                    V[rr][q] = s * vx + c * vy;
                }
                __syncthreads();
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int cc = l16 + 16 * h;
                    const double x = A[p][cc], y = A[q][cc];
                    A[p][cc] = c * x - s * y;
                    A[q][cc] = s * x + c * y;
                }
                __syncthreads();
            }
            if (!any_rot) break;
            __syncthreads();
        }
        // ---- rotate my rows ----
        for (int row = row0 + tid; row < ((a.skip & 4) ? row0 : row1); row += 256) {
            double x[FPB];
#pragma unroll
            for (int q = 0; q < FPB; q++) {
                const int col = (q < FJB) ? I * FJB + q : J * FJB + (q - FJB);
                x[q] = __ldcg(a.S + (size_t)col * a.ld + row);
            }
#pragma unroll 4
            for (int uu = 0; uu < FPB; uu++) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int q = 0; q < FPB; q += 2) { s0 += x[q] * V[q][uu]; s1 += x[q + 1] * V[q + 1][uu]; }
                const int col = (uu < FJB) ? I * FJB + uu : J * FJB + (uu - FJB);
                __stcg(a.S + (size_t)col * a.ld + row, s0 + s1);
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            *(volatile int*)(a.done + I * a.R + r) = st + 1;
            *(volatile int*)(a.done + J * a.R + r) = st + 1;
        }
    }
}

__global__ void __launch_bounds__(128)
jac_colnorm_kernel(const double* __restrict__ S, int64_t ld, int m, double* __restrict__ P) {
    __shared__ double sh[4];
    const int col = blockIdx.x;
    double s = 0.0;
    for (int r = threadIdx.x; r < m; r += 128) {
        const double t = S[(size_t)col * ld + r];
        s += t * t;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) P[col] = sh[0] + sh[1] + sh[2] + sh[3];
}

// Stopping-test threshold for a sweep.  Only the chi_max largest singular triplets are kept and the
// truncation rule needs nothing of the rest but its total weight (= trace - kept).  If the columns are split
// into a head K' and a tail with  sum(tail norms^2) < (k-th largest norm^2), then once every pair touching K'
// is orthogonal the Gram matrix is block diagonal with lambda_max(tail block) <= trace(tail block) < sigma_k^2,
// so the k largest columns ARE the k largest singular triplets whatever the state of the tail block.
// theta = smallest column weight that still belongs to K' (0 when everything must be resolved).
__global__ void __launch_bounds__(1024)
jac_theta_kernel(const double* __restrict__ P, int n, int npad, int k, double* __restrict__ Psorted_tmp,
                 double* __restrict__ theta_out) {
    for (int j = threadIdx.x; j < npad; j += blockDim.x) {
        const double pj = P[j];
        int rank = 0;
        for (int q = 0; q < npad; q++) {
            const double pq = P[q];
            rank += (pq > pj) || (pq == pj && q < j);
        }
        Psorted_tmp[rank] = pj;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double theta = 0.0;
        if (n > k) {
            const double pk = Psorted_tmp[k - 1];
            double tail = 0.0;
            int idx = n;                                   // K' = [0, idx)
            while (idx > k && tail + Psorted_tmp[idx - 1] < 0.5 * pk) { tail += Psorted_tmp[idx - 1]; idx--; }
            theta = (idx < n) ? Psorted_tmp[idx - 1] : 0.0;
        }
        *theta_out = theta;
    }
}

// rank columns by descending weight, apply the NDTensors truncation rule; single block.
// out: perm[k] = column holding the k-th largest weight, Psorted[k], iscal[0] = chi_new
__global__ void __launch_bounds__(1024)
jac_sort_trunc_kernel(const double* __restrict__ P, int n, int npad, int maxdim, double cutoff,
                      int* __restrict__ perm, double* __restrict__ Psorted, int* __restrict__ iscal) {
    for (int j = threadIdx.x; j < npad; j += blockDim.x) {
        const double pj = P[j];
        int rank = 0;
        for (int k = 0; k < npad; k++) {
            const double pk = P[k];
            rank += (pk > pj) || (pk == pj && k < j);
        }
        perm[rank] = j;
        Psorted[rank] = pj;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // NDTensors.truncate! (un-vendored, v0.3.74): P has min(m, n) entries in the reference;
        // the extra (exactly or numerically) zero weights here are removed by the same rule.
        int len = n;
        double scale = 0.0;
        for (int k = 0; k < len; k++) scale += Psorted[k];
        if (scale == 0.0) scale = 1.0;
        int keep = len;
        double err = 0.0;
        while (keep > maxdim) { err += Psorted[keep - 1]; keep--; }
        while (keep > 1 && err + Psorted[keep - 1] <= cutoff * scale) { err += Psorted[keep - 1]; keep--; }
        if (keep < 1) keep = 1;
        iscal[0] = keep;
    }
}

// label core [c][x + Dx*k] = S[c*Dx + x, perm[k]];  ortho core [y + n*k] = S[m + y, perm[k]]
__global__ void __launch_bounds__(256)
jac_gather_kernel(const double* __restrict__ S, int64_t ld, int m, int n, int C, const int* __restrict__ perm,
                  const int* __restrict__ iscal, double* __restrict__ label_core, double* __restrict__ ortho_core) {
    const int chi = iscal[0];
    const int Dx = m / C;
    const int64_t tot = (int64_t)(m + n) * chi;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e / (m + n));
        const int r = (int)(e - (int64_t)k * (m + n));
        const double v = S[(size_t)perm[k] * ld + r];
        if (r < m) {
            const int c = r / Dx, xx = r - c * Dx;
            label_core[(size_t)c * Dx * chi + xx + (size_t)Dx * k] = v;
        } else {
            ortho_core[(size_t)(r - m) + (size_t)n * k] = v;
        }
    }
}
}  // namespace

template <int PB>
static int jacobi_sweeps(mpst_ctx* c, int m, int n, int npad, int64_t ld, double cutoff, const double* trace_dev,
                         int* sweeps_out) {
    constexpr int JB = PB / 2;
    const int nb = npad / JB;                    // even
    const int npairs = nb / 2;
    const int mm = m + n;
    int rsplit = std::max(1, std::min((m + GR - 1) / GR, (2 * c->sm_count + npairs - 1) / npairs));
    TRY(ensure_buf(c, &c->gpart, &c->gpartcap, (size_t)npairs * rsplit * PB * PB));
    TRY(ensure_buf(c, &c->wbuf, &c->wbufcap, (size_t)npairs * PB * PB));
    unsigned long long* maxoff = reinterpret_cast<unsigned long long*>(c->scal + 8);
    const double eps = 2.220446049250313e-16;
    const double tol = 1e-15;
    // The measure is taken BEFORE a sweep's rotations.  Jacobi converges quadratically, so a sweep that starts
    // with every relative off-diagonal <= 1e-8 ends with them <= ~1e-16: stop after it, no verification sweep.
    const double conv = 1e-8;
    const double abs_tol = std::max(1e-30, std::min(0.5 * eps, 1e-6 * cutoff));
    const int inner = c->flag[F_SVD_INNER];
    const size_t solve_smem = 2 * sizeof(double) * PB * (PB + 1);
    CUDA_TRY(c, cudaFuncSetAttribute(jac_solve_kernel<PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem));
    int sweeps = 0;
    const int max_sweeps = 60;
    bool converged = false;
    std::string hist;
    for (; sweeps < max_sweeps && !converged; sweeps++) {
        CUDA_TRY(c, cudaMemsetAsync(maxoff, 0, sizeof(unsigned long long), c->stream));
        for (int st = 0; st < nb - 1; st++) {
            jac_gram_kernel<PB><<<dim3(npairs, rsplit), 256, 0, c->stream>>>(c->S, ld, m, nb, st, rsplit, c->gpart);
            jac_solve_kernel<PB><<<npairs, PB * 8, solve_smem, c->stream>>>(c->gpart, rsplit, c->wbuf, tol, abs_tol, trace_dev, inner, maxoff);
            jac_apply_kernel<PB><<<dim3(npairs, (mm + 127) / 128), 128, 0, c->stream>>>(c->S, ld, mm, nb, st, c->wbuf);
            c->launches += 3;
        }
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaMemcpyAsync(c->hscal + 8, maxoff, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (c->hscal[8] <= conv) converged = true;
        if (sweeps < 64) { char b[32]; snprintf(b, sizeof b, " %.2e", c->hscal[8]); hist += b; }
        if (!(c->hscal[8] == c->hscal[8])) break;                 // NaN in the bond tensor
    }
    if (c->flag[F_SVD_DEBUG]) fprintf(stderr, "[svd] m=%d n=%d PB=%d sweeps=%d:%s\n", m, n, PB, sweeps, hist.c_str());
    if (!converged) {
        char b[160];
        snprintf(b, sizeof b, "svd: Jacobi did not converge (m=%d n=%d conv=%.2e) max-offdiag per sweep:", m, n, conv);
        c->err = std::string(b) + hist;
        return MPST_E_NUMERIC;
    }
    if (sweeps_out) *sweeps_out = sweeps;
    return MPST_OK;
}

static int jacobi_sweeps_fused(mpst_ctx* c, int m, int n, int npad, int64_t ld, double cutoff, int chi_max,
                               const double* trace_dev, int* sweeps_out, bool* used) {
    *used = false;
    const int nb = npad / FJB, npairs = nb / 2, mm = m + n;
    int occ = 0;
    CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jac_sweep_fused_kernel, 256, 0));
    const int cap = occ * c->sm_count;
    if (cap < npairs) return MPST_OK;                          // cannot co-schedule: caller uses the 3-kernel path
    int R = std::min(cap / npairs, (mm + GR - 1) / GR);
    R = std::max(1, std::min(R, 64));
    const int rows_per = (int)round_up((mm + R - 1) / R, 8);
    R = (mm + rows_per - 1) / rows_per;
    TRY(ensure_buf(c, &c->gpart, &c->gpartcap, 2 * (size_t)npairs * R * FPB * FPB));
    const size_t nflags = (size_t)nb * R + npairs + 8;
    if (nflags > c->flagcap) {
        if (c->flags) cudaFree(c->flags);
        CUDA_TRY(c, cudaMalloc(&c->flags, nflags * 2 * sizeof(int)));
        c->flagcap = nflags * 2;
    }
    FusedArgs a;
    a.S = c->S; a.ld = ld; a.m = m; a.mm = mm; a.nb = nb; a.R = R; a.rows_per = rows_per; a.gpart = c->gpart;
    a.done = c->flags; a.garr = c->flags + (size_t)nb * R; a.errflag = a.garr + npairs;
    a.maxoff = reinterpret_cast<unsigned long long*>(c->scal + 8);
    a.tol = 1e-15;
    const double eps = 2.220446049250313e-16;
    a.abs_tol = std::max(1e-30, std::min(0.5 * eps, 1e-6 * cutoff));
    a.trace_dev = trace_dev;
    a.inner = c->flag[F_SVD_INNER];
    a.skip = c->flag[F_SVD_SKIP];
    const int fixed = c->flag[F_SVD_FIXED];
    const bool partial = !c->flag[F_SVD_FULL];
    a.theta = c->scal + 6;
    CUDA_TRY(c, cudaMemsetAsync(c->scal + 6, 0, sizeof(double), c->stream));
    const double conv = 1e-8;
    int sweeps = 0;
    bool converged = false;
    std::string hist;
    const auto t0 = std::chrono::steady_clock::now();
    for (; sweeps < 60 && !converged; sweeps++) {
        if (fixed && sweeps >= fixed) { converged = true; break; }
        if (partial) {
            jac_colnorm_kernel<<<npad, 128, 0, c->stream>>>(c->S, ld, m, c->colnorm);
            jac_theta_kernel<<<1, 1024, 0, c->stream>>>(c->colnorm, n, npad, chi_max, c->colnorm + npad, c->scal + 6);
            c->launches += 2;
        }
        CUDA_TRY(c, cudaMemsetAsync(c->flags, 0, nflags * sizeof(int), c->stream));
        CUDA_TRY(c, cudaMemsetAsync(a.maxoff, 0, sizeof(unsigned long long), c->stream));
        void* args[] = {&a};
        CUDA_TRY(c, cudaLaunchCooperativeKernel((void*)jac_sweep_fused_kernel, dim3(npairs * R), dim3(256), args, 0, c->stream));
        c->launches++;
        CUDA_TRY(c, cudaMemcpyAsync(c->hscal + 8, a.maxoff, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->hiscal + 4, a.errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (c->hiscal[4]) { c->err = "svd: fused Jacobi sweep timed out on a dependency wait"; return MPST_E_NUMERIC; }
        if (c->hscal[8] <= conv && !fixed) converged = true;
        if (sweeps < 64) { char b[32]; snprintf(b, sizeof b, " %.2e", c->hscal[8]); hist += b; }
        if (!(c->hscal[8] == c->hscal[8])) break;
    }
    if (c->flag[F_SVD_DEBUG]) fprintf(stderr, "[svd fused] m=%d n=%d R=%d sweeps=%d t=%.3f ms:%s\n", m, n, R, sweeps,
                                          1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), hist.c_str());
    if (!converged) {
        char b[160];
        snprintf(b, sizeof b, "svd: Jacobi did not converge (m=%d n=%d) max-offdiag per sweep:", m, n);
        c->err = std::string(b) + hist;
        return MPST_E_NUMERIC;
    }
    if (sweeps_out) *sweeps_out = sweeps;
    *used = true;
    return MPST_OK;
}

// Called while the gradient kernel of the same bond is still running: builds (first time) and uploads the CUDA graph
// of the split's first subspace round, so that the split itself starts with a single cheap launch.  Must see the same
// arguments as the svd_split_device call that follows; changes no numerical state.
int svd_split_prepare(mpst_ctx* c, int Dl, int Dr, int C, int going_left, int chi_max, double cutoff, const double* norm2_dev) {
    if (c->flag[F_SVD_NOPREP] || c->flag[F_SVD_NOGRAPH] || c->flag[F_SVD_DEBUG] || c->flag[F_SVD_NOSUB] || !c->stream2) return MPST_OK;
    const int n = going_left ? Dr : Dl;
    const int m = C * (going_left ? Dl : Dr);
    const bool wide = c->flag[F_SVD_PB64] && n >= 256;
    const int npad = (int)round_up(n, wide ? 64 : 32);
    const int64_t ld = round_up(m + n, 2);
    TRY(ensure_buf(c, &c->S, &c->Scap, (size_t)ld * npad));
    const double* trace_dev = norm2_dev ? nullptr : c->scal + 5;
    bool done = false;
    int chi_dummy = 0;
    c->svd_prepare_only = true;
    const int rc = svd_subspace_device(c, c->S, ld, m, n, C, chi_max, cutoff, trace_dev, nullptr, nullptr, &chi_dummy, nullptr, &done);
    c->svd_prepare_only = false;
    return rc;
}

// B: [C][Dl*Dr] on the device.  Writes the two new cores into label_core / ortho_core (device,
// capacity checked by the caller) and returns chi_new (host) after one small D2H copy.
// norm2_dev != nullptr: B is scaled by 1/sqrt(*norm2_dev) on load (fused renormalisation).
int svd_split_device(mpst_ctx* c, const double* B, int Dl, int Dr, int C, int going_left, int chi_max,
                     double cutoff, const double* norm2_dev, double* label_core, double* ortho_core,
                     int* chi_new, double* sigma_host, int* sweeps_out) {
    const int n = going_left ? Dr : Dl;
    const int m = C * (going_left ? Dl : Dr);
    const bool wide = c->flag[F_SVD_PB64] && n >= 256;
    const int PBsel = wide ? 64 : 32;
    const int npad = (int)round_up(n, PBsel);
    const int mm = m + n;
    const int64_t ld = round_up(mm, 2);
    TRY(ensure_buf(c, &c->S, &c->Scap, (size_t)ld * npad));
    {
        const int64_t tot = (int64_t)mm * npad;
        jac_load_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(B, c->S, going_left, Dl, Dr, C, m, n, npad,
                                                                             ld, norm2_dev, norm2_dev != nullptr);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
    }
    // ||M||_F^2 for the absolute rotation floor: 1 when the load normalises, else computed here
    const double* trace_dev = nullptr;
    if (!norm2_dev) {
        TRY(launch_sumsq(c, B, (int64_t)Dl * Dr * C, c->scal + 5));
        trace_dev = c->scal + 5;
    }
    c->last[L_SVD_CALLS]++;
    {   // fast path: subspace iteration + Rayleigh-Ritz (svd_subspace.cu); falls through when not applicable
        bool done = false;
        TRY(svd_subspace_device(c, c->S, ld, m, n, C, chi_max, cutoff, trace_dev, label_core, ortho_core, chi_new, sigma_host, &done));
        if (done) return MPST_OK;
    }
    bool fused = false;
    if (!wide && !c->flag[F_SVD_LEGACY]) TRY(jacobi_sweeps_fused(c, m, n, npad, ld, cutoff, chi_max, trace_dev, sweeps_out, &fused));
    if (!fused) {
        if (wide) TRY(jacobi_sweeps<64>(c, m, n, npad, ld, cutoff, trace_dev, sweeps_out));
        else TRY(jacobi_sweeps<32>(c, m, n, npad, ld, cutoff, trace_dev, sweeps_out));
    }
    c->last[L_SVD_PATH] = fused ? 4 : 5;
    c->last[L_SVD_ITERS] = 0;
    c->last[L_SVD_JACOBI]++;
    jac_colnorm_kernel<<<npad, 128, 0, c->stream>>>(c->S, ld, m, c->colnorm);
    jac_sort_trunc_kernel<<<1, 1024, 0, c->stream>>>(c->colnorm, n, npad, chi_max, cutoff, c->perm, c->colnorm + npad, c->iscal);
    jac_gather_kernel<<<2 * c->sm_count, 256, 0, c->stream>>>(c->S, ld, m, n, C, c->perm, c->iscal, label_core, ortho_core);
    c->launches += 3;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(c->hiscal, c->iscal, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));   // chi_new, non-finite flag
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *chi_new = c->hiscal[0];
    if (sigma_host) {
        std::vector<double> tmp(*chi_new);
        CUDA_TRY(c, cudaMemcpy(tmp.data(), c->colnorm + npad, sizeof(double) * (*chi_new), cudaMemcpyDeviceToHost));
        for (int k = 0; k < *chi_new; k++) sigma_host[k] = sqrt(tmp[k]);
    }
    return MPST_OK;
}
