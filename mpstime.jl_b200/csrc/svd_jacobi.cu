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
#include "mpst_common.cuh"

namespace {
constexpr int JB = 16, PB = 2 * JB;      // block width, pair width
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

__global__ void __launch_bounds__(256)
jac_gram_kernel(const double* __restrict__ S, int64_t ld, int m, int nb, int st, int rsplit,
                double* __restrict__ gpart) {
    __shared__ double T[PB][GR + 1];
    int I, J;
    rr_pair(nb, st, blockIdx.x, I, J);
    const int split = blockIdx.y;
    const int nchunks = (m + GR - 1) / GR;
    const int c0 = (int)(((int64_t)nchunks * split) / rsplit), c1 = (int)(((int64_t)nchunks * (split + 1)) / rsplit);
    const int tid = threadIdx.x;
    const int u = tid & 15, v = tid >> 4;
    double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
    const int lr = tid & 63, lg = tid >> 6;                 // loader: row, column group of 8
    for (int ch = c0; ch < c1; ch++) {
        const int r = ch * GR + lr;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int cc = lg * 8 + k;
            const int col = (cc < JB) ? I * JB + cc : J * JB + (cc - JB);
            T[cc][lr] = (r < m) ? S[(size_t)col * ld + r] : 0.0;
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
    double* g = gpart + ((size_t)blockIdx.x * rsplit + split) * PB * PB;
    g[(2 * u) * PB + 2 * v] = a00;
    g[(2 * u) * PB + 2 * v + 1] = a01;
    g[(2 * u + 1) * PB + 2 * v] = a10;
    g[(2 * u + 1) * PB + 2 * v + 1] = a11;
}

// eigen-decomposition of the pair Gram matrix; W[v][u] = component v of eigenvector u
__global__ void __launch_bounds__(256)
jac_solve_kernel(const double* __restrict__ gpart, int rsplit, double* __restrict__ wbuf, double tol,
                 double negl, unsigned long long* __restrict__ maxoff_bits) {
    __shared__ double A[PB][PB + 1];
    __shared__ double V[PB][PB + 1];
    __shared__ double cs[JB][2];
    __shared__ int any_rot;
    const int tid = threadIdx.x;
    const double* g = gpart + (size_t)blockIdx.x * rsplit * PB * PB;
    for (int e = tid; e < PB * PB; e += 256) {
        double s = 0.0;
        for (int k = 0; k < rsplit; k++) s += g[(size_t)k * PB * PB + e];
        A[e / PB][e % PB] = s;
        V[e / PB][e % PB] = (e / PB == e % PB) ? 1.0 : 0.0;
    }
    __syncthreads();
    // symmetrise (partials are computed as full blocks; keep exact symmetry) + convergence measure
    double mo = 0.0;
    for (int e = tid; e < PB * PB; e += 256) {
        const int r = e / PB, cidx = e % PB;
        if (r < cidx) {
            const double apq = 0.5 * (A[r][cidx] + A[cidx][r]);
            const double den = A[r][r] * A[cidx][cidx];
            if (den > negl * negl && apq != 0.0) mo = fmax(mo, fabs(apq) / sqrt(den));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mo = fmax(mo, __shfl_xor_sync(0xffffffffu, mo, o));
    if ((tid & 31) == 0 && mo > 0.0) atomicMax(maxoff_bits, (unsigned long long)__double_as_longlong(mo));
    __syncthreads();

    const int k = tid >> 4, l16 = tid & 15;
    for (int sweep = 0; sweep < 40; sweep++) {
        if (tid == 0) any_rot = 0;
        __syncthreads();
        for (int rd = 0; rd < PB - 1; rd++) {
            int p, q;
            rr_pair(PB, rd, k, p, q);
            const double app = A[p][p], aqq = A[q][q], apq = A[p][q];
            double c = 1.0, s = 0.0;
            if (apq != 0.0 && fabs(apq) > tol * sqrt(fabs(app * aqq))) {
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                c = 1.0 / sqrt(1.0 + t * t);
                s = t * c;
                if (l16 == 0) any_rot = 1;
            }
            __syncthreads();
#pragma unroll
            for (int h = 0; h < 2; h++) {               // columns p, q of A and V
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
            for (int h = 0; h < 2; h++) {               // rows p, q of A
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
    for (int e = tid; e < PB * PB; e += 256) w[e] = V[e / PB][e % PB];
}

__global__ void __launch_bounds__(256)
jac_apply_kernel(double* __restrict__ S, int64_t ld, int mm, int nb, int st, const double* __restrict__ wbuf) {
    __shared__ double W[PB][PB];
    int I, J;
    rr_pair(nb, st, blockIdx.x, I, J);
    const double* w = wbuf + (size_t)blockIdx.x * PB * PB;
    for (int e = threadIdx.x; e < PB * PB; e += 256) W[e / PB][e % PB] = w[e];
    __syncthreads();
    const int r = blockIdx.y * 256 + threadIdx.x;
    if (r >= mm) return;
    double x[PB];
#pragma unroll
    for (int v = 0; v < PB; v++) {
        const int col = (v < JB) ? I * JB + v : J * JB + (v - JB);
        x[v] = S[(size_t)col * ld + r];
    }
#pragma unroll 4
    for (int u = 0; u < PB; u++) {
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < PB; v++) s += x[v] * W[v][u];
        const int col = (u < JB) ? I * JB + u : J * JB + (u - JB);
        S[(size_t)col * ld + r] = s;
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

// B: [C][Dl*Dr] on the device.  Writes the two new cores into label_core / ortho_core (device,
// capacity checked by the caller) and returns chi_new (host) after one small D2H copy.
// norm2_dev != nullptr: B is scaled by 1/sqrt(*norm2_dev) on load (fused renormalisation).
int svd_split_device(mpst_ctx* c, const double* B, int Dl, int Dr, int C, int going_left, int chi_max,
                     double cutoff, const double* norm2_dev, double* label_core, double* ortho_core,
                     int* chi_new, double* sigma_host, int* sweeps_out) {
    const int n = going_left ? Dr : Dl;
    const int m = C * (going_left ? Dl : Dr);
    const int npad = (int)round_up(n, PB);
    const int nb = npad / JB;                    // even
    const int npairs = nb / 2;
    const int mm = m + n;
    const int64_t ld = round_up(mm, 2);
    TRY(ensure_buf(c, &c->S, &c->Scap, (size_t)ld * npad));
    int rsplit = std::max(1, std::min((m + GR - 1) / GR, (2 * c->sm_count + npairs - 1) / npairs));
    TRY(ensure_buf(c, &c->gpart, &c->gpartcap, (size_t)npairs * rsplit * PB * PB));
    TRY(ensure_buf(c, &c->wbuf, &c->wbufcap, (size_t)npairs * PB * PB));
    {
        const int64_t tot = (int64_t)mm * npad;
        jac_load_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(B, c->S, going_left, Dl, Dr, C, m, n, npad,
                                                                             ld, norm2_dev, norm2_dev != nullptr);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
    }
    unsigned long long* maxoff = reinterpret_cast<unsigned long long*>(c->scal + 8);
    const double tol = 1e-15, negl = 1e-26;
    const double conv = 8.0 * 2.220446049250313e-16 * sqrt((double)m);   // LAPACK dgesvj-style sqrt(m)*eps
    int sweeps = 0;
    const int max_sweeps = 40;
    bool converged = false;
    for (; sweeps < max_sweeps && !converged; sweeps++) {
        CUDA_TRY(c, cudaMemsetAsync(maxoff, 0, sizeof(unsigned long long), c->stream));
        for (int st = 0; st < nb - 1; st++) {
            jac_gram_kernel<<<dim3(npairs, rsplit), 256, 0, c->stream>>>(c->S, ld, m, nb, st, rsplit, c->gpart);
            jac_solve_kernel<<<npairs, 256, 0, c->stream>>>(c->gpart, rsplit, c->wbuf, tol, negl, maxoff);
            jac_apply_kernel<<<dim3(npairs, (mm + 255) / 256), 256, 0, c->stream>>>(c->S, ld, mm, nb, st, c->wbuf);
            c->launches += 3;
        }
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaMemcpyAsync(c->hscal + 8, maxoff, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (c->hscal[8] <= conv) converged = true;
    }
    if (!converged) { c->err = "svd: Jacobi did not converge"; return MPST_E_NUMERIC; }
    if (sweeps_out) *sweeps_out = sweeps;
    jac_colnorm_kernel<<<npad, 128, 0, c->stream>>>(c->S, ld, m, c->colnorm);
    jac_sort_trunc_kernel<<<1, 1024, 0, c->stream>>>(c->colnorm, n, npad, chi_max, cutoff, c->perm, c->colnorm + npad, c->iscal);
    jac_gather_kernel<<<2 * c->sm_count, 256, 0, c->stream>>>(c->S, ld, m, n, C, c->perm, c->iscal, label_core, ortho_core);
    c->launches += 3;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(c->hiscal, c->iscal, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *chi_new = c->hiscal[0];
    if (sigma_host) {
        std::vector<double> tmp(*chi_new);
        CUDA_TRY(c, cudaMemcpy(tmp.data(), c->colnorm + npad, sizeof(double) * (*chi_new), cudaMemcpyDeviceToHost));
        for (int k = 0; k < *chi_new; k++) sigma_host[k] = sqrt(tmp[k]);
    }
    return MPST_OK;
}
