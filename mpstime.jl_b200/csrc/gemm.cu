// Plain FP64 tensor-core GEMM (DMMA m8n8k4), column-major:  C[M x N] = op(A) * op(B).
// Used by the subspace SVD (K5) for its tall-skinny products (M*Q, M^T*Z, Gram matrices, basis changes);
// the hot bond GEMMs have their own fused kernels (krao_gemm.cu, bond_grad.cu).
// CTA tile 64x64, K chunks of 16, 8 warps (4x2, warp tile 16x32), register-prefetched double buffering,
// operand pitches == 4 (mod 16) doubles -> conflict-free fragment loads.
#include <algorithm>
#include "mpst_common.cuh"
#include "dmma.cuh"

namespace {
constexpr int BM = 64, BN = 64, BK = 16, LDS = BK + 4;

// A(i,k) = A[i*sai + k*sak],  B(k,j) = B[k*sbk + j*sbj].  blockIdx.z = K split: split z handles the K chunks
// [z*kper, (z+1)*kper) and writes its partial product to C + z*cz (summed by the consumer in fixed order).
__global__ void __launch_bounds__(256)
dgemm_kernel(const double* __restrict__ A, int64_t sai, int64_t sak, const double* __restrict__ B, int64_t sbk,
             int64_t sbj, double* __restrict__ C, int64_t ldc, int M, int N, int K, int kper, int64_t cz) {
    __shared__ double As[2][BM][LDS];
    __shared__ double Bs[2][BN][LDS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
    const bool a_i_fast = sai == 1, b_k_fast = sbk == 1;
    double ra[4], rb[4];
    auto load = [&](int kc) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            int i, k;
            if (a_i_fast) { i = tid & 63; k = (tid >> 6) + 4 * r; }
            else { k = tid & 15; i = (tid >> 4) + 16 * r; }
            const int gi = i0 + i, gk = kc * BK + k;
            ra[r] = (gi < M && gk < K) ? A[gi * sai + gk * sak] : 0.0;
            int j, kb;
            if (b_k_fast) { kb = tid & 15; j = (tid >> 4) + 16 * r; }
            else { j = tid & 63; kb = (tid >> 6) + 4 * r; }
            const int gj = j0 + j, gkb = kc * BK + kb;
            rb[r] = (gj < N && gkb < K) ? B[gkb * sbk + gj * sbj] : 0.0;
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            int i, k;
            if (a_i_fast) { i = tid & 63; k = (tid >> 6) + 4 * r; }
            else { k = tid & 15; i = (tid >> 4) + 16 * r; }
            As[buf][i][k] = ra[r];
            int j, kb;
            if (b_k_fast) { kb = tid & 15; j = (tid >> 4) + 16 * r; }
            else { j = tid & 63; kb = (tid >> 6) + 4 * r; }
            Bs[buf][j][kb] = rb[r];
        }
    };
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int nk_all = (K + BK - 1) / BK;
    const int kc0 = blockIdx.z * kper;
    const int nk = min(nk_all, kc0 + kper);
    C += (int64_t)blockIdx.z * cz;
    load(kc0);
    store(kc0 & 1);
    __syncthreads();
    const int fr = lane >> 2, fc = lane & 3;
    for (int kc = kc0; kc < nk; kc++) {
        const int cur = kc & 1;
        if (kc + 1 < nk) load(kc + 1);
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; k4++) {
            double a[2], b[4];
#pragma unroll
            for (int mi = 0; mi < 2; mi++) a[mi] = As[cur][wm * 16 + mi * 8 + fr][k4 * 4 + fc];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) b[ni] = Bs[cur][wn * 32 + ni * 8 + fr][k4 * 4 + fc];
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        if (kc + 1 < nk) store(cur ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 2; mi++) {
        const int i = i0 + wm * 16 + mi * 8 + fr;
        if (i >= M) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int j = j0 + wn * 32 + ni * 8 + 2 * fc;
            if (j < N) C[i + ldc * (int64_t)j] = acc[mi][ni][0];
            if (j + 1 < N) C[i + ldc * (int64_t)(j + 1)] = acc[mi][ni][1];
        }
    }
}

// C[i + ldc*j] = sum_z parts[z][i + M*j]   (fixed order)
__global__ void splitk_reduce_kernel(const double* __restrict__ parts, int splits, int M, int N, double* __restrict__ C,
                                     int64_t ldc) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)M * N) return;
    double v = 0.0;
    for (int z = 0; z < splits; z++) v += parts[(size_t)z * M * N + e];
    const int j = (int)(e / M), i = (int)(e - (int64_t)j * M);
    C[i + ldc * j] = v;
}
}  // namespace

// C = op(A) * op(B); ta/tb: 0 = as stored (column-major), 1 = transposed.
// The subspace SVD's products are tiny next to the machine (tens of MFLOP) but long in K: when the output tiles
// cover less than half of the SMs the K loop is split across CTAs and the partials are summed in a fixed order.
// lane 0: the context's main stream and split-K workspace; lane 1: the side stream (own workspace), used by the
// subspace SVD to run the small Gram / Cholesky chain next to the big product of the same iteration.
int launch_dgemm_on(mpst_ctx* c, int lane, int ta, int tb, int M, int N, int K, const double* A, int64_t lda,
                    const double* B, int64_t ldb, double* C, int64_t ldc) {
    if (M <= 0 || N <= 0) return MPST_OK;
    cudaStream_t st = lane ? c->stream2 : c->stream;
    double** ws = lane ? &c->gws2 : &c->gws;
    size_t* wscap = lane ? &c->gws2cap : &c->gwscap;
    const int64_t sai = ta ? lda : 1, sak = ta ? 1 : lda;
    const int64_t sbk = tb ? ldb : 1, sbj = tb ? 1 : ldb;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    const int tiles = grid.x * grid.y, nk = (K + BK - 1) / BK;
    int splits = std::min(c->sm_count / std::max(tiles, 1), nk / 4);
    if (splits >= 2) {
        const int kper = (nk + splits - 1) / splits;
        splits = (nk + kper - 1) / kper;
        if (ensure_buf(c, ws, wscap, (size_t)splits * M * N) != MPST_OK) return MPST_E_CUDA;
        grid.z = splits;
        dgemm_kernel<<<grid, 256, 0, st>>>(A, sai, sak, B, sbk, sbj, *ws, M, M, N, K, kper, (int64_t)M * N);
        splitk_reduce_kernel<<<(unsigned)(((int64_t)M * N + 255) / 256), 256, 0, st>>>(*ws, splits, M, N, C, ldc);
        c->launches += 2;
        CUDA_TRY(c, cudaGetLastError());
        return MPST_OK;
    }
    dgemm_kernel<<<grid, 256, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, nk, 0);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_dgemm(mpst_ctx* c, int ta, int tb, int M, int N, int K, const double* A, int64_t lda, const double* B,
                 int64_t ldb, double* C, int64_t ldc) {
    return launch_dgemm_on(c, 0, ta, tb, M, N, K, A, lda, B, ldb, C, ldc);
}

// split-K variant: `splits` partial products, partial z at C + z*M*N (ldc = M).  Returns the split count used.
int launch_dgemm_splitk(mpst_ctx* c, int ta, int tb, int M, int N, int K, const double* A, int64_t lda, const double* B,
                        int64_t ldb, double* Cparts, int max_splits, int* splits_out) {
    const int64_t sai = ta ? lda : 1, sak = ta ? 1 : lda;
    const int64_t sbk = tb ? ldb : 1, sbj = tb ? 1 : ldb;
    const int nk = (K + BK - 1) / BK;
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int splits = std::max(1, std::min(std::min(max_splits, nk), (2 * c->sm_count) / std::max(tiles, 1)));
    const int kper = (nk + splits - 1) / splits;
    splits = (nk + kper - 1) / kper;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, splits);
    dgemm_kernel<<<grid, 256, 0, c->stream>>>(A, sai, sak, B, sbk, sbj, Cparts, M, M, N, K, kper, (int64_t)M * N);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    *splits_out = splits;
    return MPST_OK;
}
