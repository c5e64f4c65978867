// K6 / K2-fwd for wide outputs (64 columns per block; chi = 64 at the north-star shape):
//     out_i[n] = sum_{s,a} x_i[s] E_i[a] W[s + d a][n]
// the register-operand scheme of krao_reg_kernel with W STREAMED instead of resident.  d*chi x 64 doubles (512 KB at
// d = 16, chi = 64) do not fit in shared memory, so krao_gemm_kernel builds the Khatri-Rao operand tile in shared
// memory (2 loads + multiply + store per element, a block barrier per chunk: 0.62 of the DGEMM rate).  Here
//   * W is re-packed once per launch into K-slabs of 4 link values (16 DQ rows of k = s + d a, DQ = d / 4), each slab a
//     contiguous [64 columns][16 DQ + 4] block (pitch == 4 mod 16: conflict-free fragment loads) that ONE bulk-TMA copy
//     brings into a 4-stage mbarrier ring shared by the CTA's 8 warps;
//   * a warp owns 8 MI samples for a whole pass over K: its MI x 8 accumulator fragments stay in registers, the site
//     values x_i[s] it needs (DQ per row: s = 4 q + lane%4) are loaded once per tile into registers, the environment
//     values E_i[a] come straight from global memory (one load per 4 k-steps and row, prefetched a slab ahead), and an
//     A fragment is ONE multiply: no operand tile, no block barrier.
// Per k-step: MI DMUL + 8 LDS + 8 MI DMMA (krao_gemm: ~1 non-DMMA instruction per DMMA; here 0.4 at MI = 4).  On
// sm_100 every non-DMMA instruction of a sub-partition delays its tensor pipe, so this ratio is the efficiency.
#include <algorithm>
#include "mpst_common.cuh"
#include "dmma.cuh"

namespace {

// W2[cb][slab][n][kk] = W[(cb*64 + n) * ldw + slab*KS + kk]   (zero for n >= n_out), pitch KS + 4
__global__ void pack_w_slabs_kernel(const double* __restrict__ W, int64_t ldw, int n_out, int K, int KS, int nslab,
                                    int NCOL, double* __restrict__ W2) {
    const int pitch = KS + 4;
    const int64_t tot = (int64_t)gridDim.y * nslab * NCOL * pitch;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)nslab * NCOL * pitch;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e % pitch);
        const int n = (int)((e / pitch) % NCOL);
        const int slab = (int)(e / ((int64_t)pitch * NCOL));
        const int col = blockIdx.y * NCOL + n, k = slab * KS + kk;
        const int64_t o = (int64_t)blockIdx.y * nslab * NCOL * pitch + e;
        if (o < tot) W2[o] = (kk < KS && col < n_out && k < K) ? W[(int64_t)col * ldw + k] : 0.0;
    }
}

template <int MI, int DQ, int NI, int ST, int OCC>
__global__ void __launch_bounds__(256, OCC)
krao_slab_kernel(const double* __restrict__ x, const double* __restrict__ E, const double* __restrict__ W2,
                 double* __restrict__ out, int64_t row_begin, int64_t row_end, int d, int chi, int n_out, int64_t ldo,
                 int nslab) {
    constexpr int NCOL = 8 * NI;                                           // output columns per block
    constexpr int KS = 16 * DQ, PITCH = KS + 4, SLAB = NCOL * PITCH;      // doubles per slab
    constexpr int ROWS = 8 * MI;                                           // samples per warp
    extern __shared__ __align__(16) unsigned char smraw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + ST;
    double* ring = reinterpret_cast<double*>(smraw + 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const double* Wc = W2 + (size_t)blockIdx.y * nslab * SLAB;            // this column block's slabs
    const int col0 = blockIdx.y * NCOL;

    if (tid == 0) {
        for (int s = 0; s < ST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        mbar_fence_init();
    }
    __syncthreads();

    // tiles of 8 warps x ROWS samples, a contiguous balanced range per CTA; every tile is one pass over the nslab slabs
    const int64_t t0 = row_begin / (8 * ROWS), t1 = (row_end + 8 * ROWS - 1) / (8 * ROWS);
    const int64_t ntile = t1 - t0;
    const int64_t tb = t0 + ntile * blockIdx.x / gridDim.x, te = t0 + ntile * (blockIdx.x + 1) / gridDim.x;
    const int64_t total = (te - tb) * nslab;                               // slabs this CTA consumes
    constexpr int LOOK = ST - 2;                                           // the stage refilled was released two slabs ago
    int64_t issued = 0;
    auto issue_one = [&]() {                                               // lane 0 of warp 0
        if (issued >= total) return;
        const int s = (int)(issued % ST);
        mbar_wait(&empty[s], (uint32_t)(((issued / ST) & 1) ^ 1));
        fence_proxy_async();
        mbar_expect_tx(&full[s], (uint32_t)(SLAB * sizeof(double)));
        bulk_g2s(ring + (size_t)s * SLAB, Wc + (size_t)(issued % nslab) * SLAB, (uint32_t)(SLAB * sizeof(double)), &full[s]);
        issued++;
    };
    if (tid == 0)
        for (int k = 0; k < LOOK; k++) issue_one();

    int64_t g = 0;                                                         // slabs consumed so far
    for (int64_t tile = tb; tile < te; tile++) {
        const int64_t i0 = tile * (8 * ROWS) + (int64_t)warp * ROWS;
        // this lane's rows (clamped: rows past row_end are computed on a valid row and never stored)
        const double* er[MI];
        double xq[MI][DQ];
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            int64_t i = i0 + mi * 8 + fr;
            i = i < row_end ? i : row_end - 1;
            i = i < row_begin ? row_begin : i;
            er[mi] = E + i * chi;
#pragma unroll
            for (int q = 0; q < DQ; q++) xq[mi][q] = x[i * d + 4 * q + fc];
        }
        double acc[MI][NI][2];
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        double en[MI];                                                     // E_i[a + 1], in flight while link value a is consumed
#pragma unroll
        for (int mi = 0; mi < MI; mi++) en[mi] = er[mi][0];
        for (int slab = 0; slab < nslab; slab++, g++) {
            if (tid == 0) issue_one();                                     // slab g + LOOK
            const int st = (int)(g % ST);
            mbar_wait(&full[st], (uint32_t)((g / ST) & 1));
            const double* wp = ring + (size_t)st * SLAB + (size_t)fr * PITCH + fc;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double ec[MI];
                const int anext = min(4 * slab + j + 1, chi - 1);
#pragma unroll
                for (int mi = 0; mi < MI; mi++) { ec[mi] = en[mi]; en[mi] = er[mi][anext]; }
#pragma unroll
                for (int q = 0; q < DQ; q++) {
                    const int kk = 4 * (j * DQ + q);                       // k-step inside the slab: k = s + d a, s = 4 q + fc
                    double a[MI], b[NI];
#pragma unroll
                    for (int mi = 0; mi < MI; mi++) a[mi] = xq[mi][q] * ec[mi];
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) b[ni] = wp[(size_t)ni * 8 * PITCH + kk];
#pragma unroll
                    for (int mi = 0; mi < MI; mi++)
#pragma unroll
                        for (int ni = 0; ni < NI; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            const int64_t i = i0 + mi * 8 + fr;
            if (i < row_begin || i >= row_end) continue;
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                const int n = col0 + ni * 8 + 2 * fc;
                if (n < n_out) out[i * ldo + n] = acc[mi][ni][0];
                if (n + 1 < n_out) out[i * ldo + n + 1] = acc[mi][ni][1];
            }
        }
    }
}

template <int MI, int DQ, int NI, int ST, int OCC>
int launch_slab_t(mpst_ctx* c, const double* x, const double* E, const double* W2, double* out, int64_t row_begin,
                  int64_t row_end, int d, int chi, int n_out, int64_t ldo, int nslab, int ncb) {
    constexpr int KS = 16 * DQ;
    const size_t smem = 128 + sizeof(double) * (size_t)ST * 8 * NI * (KS + 4);
    auto kern = krao_slab_kernel<MI, DQ, NI, ST, OCC>;
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntile = (row_end + 64 * MI - 1) / (64 * MI) - row_begin / (64 * MI);
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(OCC * c->sm_count / ncb, ntile));
    kern<<<dim3(gx, ncb), 256, smem, c->stream>>>(x, E, W2, out, row_begin, row_end, d, chi, n_out, ldo, nslab);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    c->last[L_KRAO_KERNEL] = 3;
    c->last[L_KRAO_SLAB_LAUNCHES]++;
    c->last[L_KRAO_VARIANT] = 100 * NI + 10 * MI + DQ;
    return MPST_OK;
}

template <int DQ, int NI>
int launch_slab_mi(mpst_ctx* c, int mi, const double* x, const double* E, const double* W2, double* out, int64_t row_begin,
                   int64_t row_end, int d, int chi, int n_out, int64_t ldo, int nslab, int ncb) {
#define SLAB_ARGS c, x, E, W2, out, row_begin, row_end, d, chi, n_out, ldo, nslab, ncb
    if constexpr (NI == 8) {
        if (mi == 4) return launch_slab_t<4, DQ, NI, 4, 1>(SLAB_ARGS);
        if (mi == 22) return launch_slab_t<2, DQ, NI, 3, 2>(SLAB_ARGS);     // two CTAs per SM: 3-stage ring, 128 registers
    }
    if constexpr (NI <= 4) return launch_slab_t<4, DQ, NI, 4, 1>(SLAB_ARGS); // narrow blocks: 32 samples per warp
    return launch_slab_t<2, DQ, NI, 4, 1>(SLAB_ARGS);
#undef SLAB_ARGS
}

template <int DQ>
int launch_slab_ni(mpst_ctx* c, int ni, int mi, const double* x, const double* E, const double* W2, double* out,
                   int64_t row_begin, int64_t row_end, int d, int chi, int n_out, int64_t ldo, int nslab, int ncb) {
#define SLAB_ARGS c, mi, x, E, W2, out, row_begin, row_end, d, chi, n_out, ldo, nslab, ncb
    switch (ni) {
        case 4: return launch_slab_mi<DQ, 4>(SLAB_ARGS);
        case 5: return launch_slab_mi<DQ, 5>(SLAB_ARGS);
        case 6: return launch_slab_mi<DQ, 6>(SLAB_ARGS);
        case 7: return launch_slab_mi<DQ, 7>(SLAB_ARGS);
        default: return launch_slab_mi<DQ, 8>(SLAB_ARGS);
    }
#undef SLAB_ARGS
}
}  // namespace

// *handled = false: shape not covered (caller uses krao_reg_kernel / krao_gemm_kernel).  Covered: d in {8, 12, 16, 24},
// chi % 4 == 0, more than 24 output columns (split into equal blocks of at most 64), at least two tiles of rows.
int launch_krao_slab(mpst_ctx* c, const double* x, const double* E, const double* W, double* out, int64_t row_begin,
                     int64_t row_end, int d, int chi, int n_out, int64_t ldw, int64_t ldo, bool* handled) {
    *handled = false;
    if (c->flag[F_KRAO_NOSLAB] || (d != 8 && d != 12 && d != 16 && d != 24) || (chi & 3) || n_out <= 24 ||
        n_out < c->flag[F_KRAO_SLAB_MIN] || row_end - row_begin < 512)
        return MPST_OK;
    const int DQ = d / 4, KS = 16 * DQ, nslab = chi / 4;
    const int ncb = (n_out + 63) / 64;
    const int ni = (((n_out + ncb - 1) / ncb) + 7) / 8;                     // fragments per column block, 4..8
    const int ncol = 8 * std::max(ni, 4);
    const size_t need = (size_t)ncb * nslab * ncol * (KS + 4);
    TRY(ensure_buf(c, &c->kslab, &c->kslabcap, need));
    pack_w_slabs_kernel<<<dim3(32, ncb), 256, 0, c->stream>>>(W, ldw, n_out, d * chi, KS, nslab, ncol, c->kslab);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    *handled = true;
    const int mi = c->flag[F_KRAO_SLAB_MI];
#define SLAB_ARGS c, std::max(ni, 4), mi, x, E, c->kslab, out, row_begin, row_end, d, chi, n_out, ldo, nslab, ncb
    switch (DQ) {
        case 2: return launch_slab_ni<2>(SLAB_ARGS);
        case 3: return launch_slab_ni<3>(SLAB_ARGS);
        case 4: return launch_slab_ni<4>(SLAB_ARGS);
        default: return launch_slab_ni<6>(SLAB_ARGS);
    }
#undef SLAB_ARGS
}
