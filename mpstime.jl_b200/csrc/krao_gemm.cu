// K6 (and the building block of K2-forward and K7): fused Khatri-Rao x matrix product
//     out[i][n] = sum_{s,a} x[i][s] * E[i][a] * W[(s + d*a) + ldw*n]
// i.e. the per-sample environment update of update_caches!/construct_caches
// (reference Training/RealRealHighDimension.jl:45-144): the row operand (x_i (x) E_i) is never
// materialised in HBM, it is formed chunk by chunk in shared memory from the staged x / E tiles.
//
// FP64 tensor-core GEMM (DMMA m8n8k4): CTA tile 128 samples x TN outputs, K = d*chi in chunks of
// 16.  The 128-row E and x tiles are contiguous in HBM ([N][chi] / [N][d] row-major) and arrive
// with two 1-D bulk-TMA copies on an mbarrier; W chunks are register-prefetched one chunk ahead.
// Shared-memory operand pitches are == 4 (mod 16) doubles so every DMMA fragment load is
// bank-conflict free.
#include "mpst_common.cuh"
#include "dmma.cuh"

namespace {
constexpr int KC = 16, LDP = KC + 4;

template <int TM, int TN, int WM, int WN>
__global__ void __launch_bounds__(256, (TN <= 80 ? 2 : 1))
krao_gemm_kernel(const double* __restrict__ x, const double* __restrict__ E,
                 const double* __restrict__ W, double* __restrict__ out, int64_t row_begin,
                 int64_t row_end, int d, int chi, int n_out, int64_t ldw, int64_t ldo) {
    static_assert(WM * WN == 8, "8 warps");
    constexpr int MI = TM / (8 * WM), NI = TN / (8 * WN);
    constexpr int WPT = (TN * KC + 255) / 256;           // W elements per thread per chunk
    extern __shared__ __align__(16) unsigned char smraw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smraw);
    double* Es = reinterpret_cast<double*>(smraw + 16);
    double* xs = Es + (size_t)TM * chi;
    double* Ps = xs + (size_t)TM * d;
    double* Ws = Ps + 2 * TM * LDP;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int64_t i0 = (row_begin / TM + blockIdx.x) * TM;
    const int n0 = blockIdx.y * TN;
    const int K = d * chi;
    const int nk = (K + KC - 1) / KC;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(TM * (chi + d) * sizeof(double)));
        bulk_g2s(Es, E + i0 * chi, (uint32_t)(TM * chi * sizeof(double)), bar);
        bulk_g2s(xs, x + i0 * d, (uint32_t)(TM * d * sizeof(double)), bar);
    }

    double wreg[WPT];
    auto load_w = [&](int kc) {
#pragma unroll
        for (int r = 0; r < WPT; r++) {
            int e = tid + r * 256;
            int kk = e % KC, n = e / KC;
            int p = kc * KC + kk;
            double v = 0.0;
            if (n < TN && p < K && n0 + n < n_out) v = __ldg(W + (int64_t)p + ldw * (int64_t)(n0 + n));
            wreg[r] = v;
        }
    };
    auto store_w = [&](int buf) {
#pragma unroll
        for (int r = 0; r < WPT; r++) {
            int e = tid + r * 256;
            int kk = e % KC, n = e / KC;
            if (n < TN) Ws[buf * TN * LDP + n * LDP + kk] = wreg[r];
        }
    };
    auto build_p = [&](int kc, int buf) {
        const int kk = lane & 15, isub = lane >> 4;
        const int p = kc * KC + kk;
        const int a = p / d, s = p - a * d;
        const bool ok = p < K;
#pragma unroll
        for (int r = 0; r < TM / 16; r++) {
            int i = warp * (TM / 8) + 2 * r + isub;
            double v = ok ? xs[i * d + s] * Es[(size_t)i * chi + a] : 0.0;
            Ps[buf * TM * LDP + i * LDP + kk] = v;
        }
    };

    load_w(0);
    mbar_wait(bar, 0);
    build_p(0, 0);
    store_w(0);
    __syncthreads();

    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    const int fr = lane >> 2, fc = lane & 3;
    // Warps 0-3 stage the next chunk first and then run their DMMAs, warps 4-7 the other way round: every SM
    // sub-partition hosts one warp of each kind, so its tensor pipe has DMMAs queued while the other warp is on the
    // load/multiply/store work of the staging.
    const bool stage_first = (warp & 4) == 0;
    for (int kc = 0; kc < nk; kc++) {
        const int cur = kc & 1;
        if (kc + 1 < nk) load_w(kc + 1);
        if (stage_first && kc + 1 < nk) {
            build_p(kc + 1, cur ^ 1);
            store_w(cur ^ 1);
        }
        const double* Pc = Ps + cur * TM * LDP + (wm * (TM / WM) + fr) * LDP + fc;
        const double* Wc = Ws + cur * TN * LDP + (wn * (TN / WN) + fr) * LDP + fc;
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; k4++) {
            double a[MI], b[NI];
#pragma unroll
            for (int mi = 0; mi < MI; mi++) a[mi] = Pc[mi * 8 * LDP + k4 * 4];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) b[ni] = Wc[ni * 8 * LDP + k4 * 4];
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        if (!stage_first && kc + 1 < nk) {
            build_p(kc + 1, cur ^ 1);
            store_w(cur ^ 1);
        }
        __syncthreads();
    }

#pragma unroll
    for (int mi = 0; mi < MI; mi++) {
        const int64_t i = i0 + wm * (TM / WM) + mi * 8 + fr;
        if (i < row_begin || i >= row_end) continue;
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            const int n = n0 + wn * (TN / WN) + ni * 8 + 2 * fc;
            if (n < n_out) out[i * ldo + n] = acc[mi][ni][0];
            if (n + 1 < n_out) out[i * ldo + n + 1] = acc[mi][ni][1];
        }
    }
}

template <int TM, int TN, int WM, int WN>
int launch_cfg(mpst_ctx* c, const double* x, const double* E, const double* W, double* out,
               int64_t row_begin, int64_t row_end, int d, int chi, int n_out, int64_t ldw, int64_t ldo) {
    const size_t smem = 16 + sizeof(double) * ((size_t)TM * chi + (size_t)TM * d + 2 * TM * LDP + 2 * TN * LDP);
    auto kern = krao_gemm_kernel<TM, TN, WM, WN>;
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t t0 = row_begin / TM, t1 = (row_end + TM - 1) / TM;
    dim3 grid((unsigned)(t1 - t0), (unsigned)((n_out + TN - 1) / TN));
    kern<<<grid, 256, smem, c->stream>>>(x, E, W, out, row_begin, row_end, d, chi, n_out, ldw, ldo);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}
}  // namespace

// rows [row_begin, row_end) of x/E/out (all indexed by absolute sample row; buffers are padded to
// a multiple of 128 rows so whole tiles may be read).
int launch_krao_gemm_rows(mpst_ctx* c, const double* x, const double* E, const double* W, double* out,
                          int64_t row_begin, int64_t row_end, int d, int chi, int n_out, int64_t ldw,
                          int64_t ldo) {
    if (row_end <= row_begin || n_out <= 0) return MPST_OK;
    if (chi > MPST_MAX_CHI || d > MPST_MAX_D) { c->err = "krao_gemm: chi/d too large"; return MPST_E_UNSUPPORTED; }
    const size_t big = 16 + sizeof(double) * ((size_t)128 * chi + (size_t)128 * d + 2 * 128 * LDP +
                                              2 * (n_out > 64 ? 128 : 64) * LDP);
#define ARGS c, x, E, W, out, row_begin, row_end, d, chi, n_out, ldw, ldo
    if (big <= 227 * 1024) {
        if (n_out <= 8) return launch_cfg<128, 8, 8, 1>(ARGS);
        if (n_out <= 16) return launch_cfg<128, 16, 8, 1>(ARGS);
        if (n_out <= 32) return launch_cfg<128, 32, 8, 1>(ARGS);
        if (n_out <= 48) return launch_cfg<128, 48, 4, 2>(ARGS);       // chi = 40 (config B): 17% padding instead of 37%
        if (n_out <= 64) return launch_cfg<128, 64, 4, 2>(ARGS);
        if (n_out <= 80) return launch_cfg<128, 80, 4, 2>(ARGS);       // chi*C = 80 (config B, labelled core)
        if (n_out <= 96) return launch_cfg<128, 96, 4, 2>(ARGS);
        return launch_cfg<128, 128, 4, 2>(ARGS);
    }
    if (n_out <= 8) return launch_cfg<64, 8, 8, 1>(ARGS);
    if (n_out <= 16) return launch_cfg<64, 16, 4, 2>(ARGS);
    if (n_out <= 32) return launch_cfg<64, 32, 4, 2>(ARGS);
    if (n_out <= 64) return launch_cfg<64, 64, 2, 4>(ARGS);
    return launch_cfg<64, 128, 2, 4>(ARGS);
#undef ARGS
}

int launch_krao_gemm(mpst_ctx* c, const double* x, const double* E, const double* W, double* out,
                     int64_t N, int d, int chi, int n_out, int64_t ldw, int64_t ldo) {
    return launch_krao_gemm_rows(c, x, E, W, out, 0, N, d, chi, n_out, ldw, ldo);
}
