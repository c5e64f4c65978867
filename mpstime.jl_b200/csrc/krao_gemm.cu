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
#include <algorithm>
#include <cstdlib>
#include "mpst_common.cuh"
#include "dmma.cuh"

int launch_krao_slab(mpst_ctx* c, const double* x, const double* E, const double* W, double* out, int64_t row_begin,
                     int64_t row_end, int d, int chi, int n_out, int64_t ldw, int64_t ldo, bool* handled);

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

// ---- register-operand variant for narrow outputs (environment update, unlabelled forward factor) ------------
// The whole weight matrix W (K x n_out, K = d*chi) is resident in shared memory for the lifetime of a persistent
// CTA (one per SM; columns arrive by bulk TMA with a pitch == 4 (mod 16), conflict free); every WARP owns 16 samples
// at a time: it stages their x / E rows in its private shared-memory region, forms the Khatri-Rao A fragments in
// registers (2 loads + 1 multiply each, reused by all n blocks) and streams through the K dimension without any
// block barrier.  Work units (16 samples) are dealt round-robin to all warps of the grid.
__host__ __device__ inline int kr_pitch4(int n) { return ((n + 11) / 16) * 16 + 4; }

template <int NI>
__global__ void __launch_bounds__(256, 1)
krao_reg_kernel(const double* __restrict__ x, const double* __restrict__ E, const double* __restrict__ W,
                double* __restrict__ out, int64_t row_begin, int64_t row_end, int d, int chi, int n_out, int64_t ldw,
                int64_t ldo) {
    extern __shared__ __align__(16) unsigned char smraw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smraw);
    double* Ws = reinterpret_cast<double*>(smraw + 16);
    const int K = d * chi;
    const int ldws = kr_pitch4(K), ldx = kr_pitch4(d), lde = kr_pitch4(chi);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* xs = Ws + (size_t)(8 * NI) * ldws + (size_t)warp * 16 * (ldx + lde);   // this warp's 16 x ldx | 16 x lde
    double* es = xs + 16 * ldx;
    const int fr = lane >> 2, fc = lane & 3;

    __shared__ int next_unit;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); next_unit = 0; }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)((size_t)n_out * K * sizeof(double)));
        __syncwarp();
        for (int n = lane; n < n_out; n += 32) bulk_g2s(Ws + (size_t)n * ldws, W + ldw * (int64_t)n, (uint32_t)(K * sizeof(double)), bar);
    }
    // columns n_out .. 8 NI - 1 feed accumulators that are never stored: zero them so no NaN garbage is multiplied
    for (int e = tid; e < (8 * NI - n_out) * ldws; e += 256) Ws[(size_t)n_out * ldws + e] = 0.0;
    __syncthreads();
    mbar_wait(bar, 0);

    // 16-sample units: a contiguous, balanced range per CTA; its warps pull the next unit from a shared counter, so
    // every warp of the SM stays busy until the CTA's range is exhausted
    const int64_t u0 = row_begin / 16, u1 = (row_end + 15) / 16;
    const int64_t nun = u1 - u0;
    const int64_t cb = u0 + nun * blockIdx.x / gridDim.x, ce = u0 + nun * (blockIdx.x + 1) / gridDim.x;
    while (true) {
        int64_t u = 0;
        if (lane == 0) u = cb + atomicAdd(&next_unit, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= ce) break;
        const int64_t i0 = u * 16;
        __syncwarp();
        for (int e = lane; e < 16 * d; e += 32) { const int r = e / d, s = e - r * d; xs[r * ldx + s] = x[(i0 + r) * d + s]; }
        for (int e = lane; e < 16 * chi; e += 32) { const int r = e / chi, a = e - r * chi; es[r * lde + a] = E[(i0 + r) * chi + a]; }
        __syncwarp();
        double acc[2][NI][2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        // this lane's k index of the k-step: k = 4*step + fc = s + d*a
        int a = fc / d, sidx = fc - a * d;
        const double* xr0 = xs + fr * ldx;
        const double* xr1 = xs + (8 + fr) * ldx;
        const double* er0 = es + fr * lde;
        const double* er1 = es + (8 + fr) * lde;
        const double* wp = Ws + (size_t)fr * ldws + fc;
#pragma unroll 4
        for (int k0 = 0; k0 < K; k0 += 4) {
            const bool kin = k0 + fc < K;                                  // K is a multiple of 4 in every shipped config
            const double a0 = kin ? xr0[sidx] * er0[a] : 0.0;
            const double a1 = kin ? xr1[sidx] * er1[a] : 0.0;
            double b[NI];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) b[ni] = kin ? wp[(size_t)ni * 8 * ldws + k0] : 0.0;
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                dmma_8x8x4(acc[0][ni][0], acc[0][ni][1], a0, b[ni]);
                dmma_8x8x4(acc[1][ni][0], acc[1][ni][1], a1, b[ni]);
            }
            sidx += 4;
            while (sidx >= d) { sidx -= d; a++; }
        }
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const int64_t i = i0 + mi * 8 + fr;
            if (i < row_begin || i >= row_end) continue;
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                const int n = ni * 8 + 2 * fc;
                if (n < n_out) out[i * ldo + n] = acc[mi][ni][0];
                if (n + 1 < n_out) out[i * ldo + n + 1] = acc[mi][ni][1];
            }
        }
    }
}

template <int NI>
int launch_reg(mpst_ctx* c, const double* x, const double* E, const double* W, double* out, int64_t row_begin,
               int64_t row_end, int d, int chi, int n_out, int64_t ldw, int64_t ldo, size_t smem) {
    c->last[L_KRAO_KERNEL] = 1;
    c->last[L_KRAO_VARIANT] = NI;
    c->last[L_KRAO_REG_MASK] |= 1 << NI;
    auto kern = krao_reg_kernel<NI>;
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t units = (row_end + 15) / 16 - row_begin / 16;
    const int grid = (int)std::min<int64_t>(c->sm_count, (units + 7) / 8);
    kern<<<grid, 256, smem, c->stream>>>(x, E, W, out, row_begin, row_end, d, chi, n_out, ldw, ldo);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

template <int TM, int TN, int WM, int WN>
int launch_cfg(mpst_ctx* c, const double* x, const double* E, const double* W, double* out,
               int64_t row_begin, int64_t row_end, int d, int chi, int n_out, int64_t ldw, int64_t ldo) {
    const size_t smem = 16 + sizeof(double) * ((size_t)TM * chi + (size_t)TM * d + 2 * TM * LDP + 2 * TN * LDP);
    c->last[L_KRAO_KERNEL] = 2;
    c->last[L_KRAO_VARIANT] = TN;
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
    {   // wide outputs: W streamed in K-slabs through the register-operand scheme (krao_slab.cu)
        bool handled = false;
        TRY(launch_krao_slab(ARGS, &handled));
        if (handled) return MPST_OK;
    }
    // narrow outputs with W resident in shared memory: the register-operand kernel (needs 16-byte columns, d >= 4)
    if (n_out <= 48 && n_out >= 8 && d >= 4 && ((d * chi) % 4) == 0 && (ldw % 2) == 0 && row_end - row_begin >= 256 &&
        !c->flag[F_KRAO_NOREG]) {
        const int NIc = (n_out + 7) / 8;
        const size_t smem = 16 + sizeof(double) * ((size_t)8 * NIc * kr_pitch4(d * chi) + (size_t)8 * 16 * (kr_pitch4(d) + kr_pitch4(chi)));
        if (smem <= 227 * 1024) {
            switch (NIc) {
                case 1: return launch_reg<1>(ARGS, smem);
                case 2: return launch_reg<2>(ARGS, smem);
                case 3: return launch_reg<3>(ARGS, smem);
                case 4: return launch_reg<4>(ARGS, smem);
                case 5: return launch_reg<5>(ARGS, smem);
                default: return launch_reg<6>(ARGS, smem);
            }
        }
    }
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
