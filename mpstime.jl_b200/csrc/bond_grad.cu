// K2 (gradient half): the summed two-site gradient of Loss_Grad_KLD / Loss_Grad_MSE
// (reference Training/loss_functions.jl:322-432, 561-619; inner loop kron_scaleadd_* :203-262,
// 435-496):
//     G_c[p, q] = sum_i w_c[i] * (xl_i (x) L_i)[p] * (xr_i (x) R_i)[q]
//     p = s_l + d*a   (rows, d*chi_l),   q = s_r + d*b   (cols, d*chi_r)
// with w_c[i] = -1/(N yhat_i) on the samples of class c (KLD) or (yhat_ic - delta)/N (MSE).
// The reference forms phi~_i explicitly (D doubles per sample) and does a rank-1 update per
// sample; here it is one FP64 tensor-core GEMM  G_c = P^T diag(w_c) Q  whose K dimension is the
// sample index and whose two Khatri-Rao operands are formed in shared memory from the raw
// per-sample factors (L, R rows and the two encoded site vectors), never in HBM.
//
// Schedule: stream-K.  The (class, p-tile, q-tile, 16-sample chunk) space is cut into one
// contiguous range per CTA (grid = #SMs), so every SM gets the same number of chunks whatever
// the tile count; a CTA writes one partial tile per (tile, range) segment and a second kernel
// sums the segments of each tile in a fixed order (deterministic, no atomics).
// Per chunk the raw factors arrive by 1-D bulk TMA (rows of 16 consecutive samples are contiguous)
// on a 2-stage mbarrier ring; the 128x16 / 16x128 operand tiles are double buffered.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "mpst_common.cuh"
#include "dmma.cuh"
#include "streamk.h"

namespace {
// CTA tile TP x TQ = (32*MI) x (16*NI): 8 warps as 4 (p) x 2 (q), warp tile (8*MI) x (8*NI).
// Two tile sizes per dimension (128: MI=4/NI=8, 96: MI=3/NI=6) so that d*chi = 480 (config B) and 1024 (north
// star) are both covered without padded rows/columns -- padding is pure DMMA waste.

struct RawStage {
    double* L;
    double* R;
    double* xl;
    double* xr;
    double* w;
};

template <int KC, int MI, int NI>
__global__ void __launch_bounds__(256, 1)
bond_grad_kernel(const double* __restrict__ xl, const double* __restrict__ xr,
                 const double* __restrict__ L, const double* __restrict__ R,
                 const double* __restrict__ w, int64_t wstride, int d, int chi_l, int chi_r,
                 const GradSeg* __restrict__ segs, const int* __restrict__ cta_ptr,
                 double* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char smraw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw);          // 2 barriers
    double* base = reinterpret_cast<double*>(smraw + 16);
    const int raw_sz = KC * (chi_l + chi_r + 2 * d + 1);
    double* raw[2] = {base, base + raw_sz};
    constexpr int TP = 32 * MI, TQ = 16 * NI;
    constexpr int LDT = (TP > TQ ? TP : TQ) + 4;               // operand pitch == 4 (mod 16): conflict-free
    static_assert(LDT % 16 == 4 && ((TP == 128 && TQ == 128) || TP + TQ == 256), "tile shape");
    double* Pt = base + 2 * raw_sz;                               // [2][KC][LDT]
    double* Qt = Pt + 2 * KC * LDT;
    const int Dl = d * chi_l, Dr = d * chi_r;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wp = warp & 3, wq = warp >> 2;
    const int fr = lane >> 2, fc = lane & 3;
    const uint32_t raw_bytes = (uint32_t)(raw_sz * sizeof(double));

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};

    auto stage_ptrs = [&](int s) {
        RawStage r;
        r.L = raw[s];
        r.R = r.L + KC * chi_l;
        r.xl = r.R + KC * chi_r;
        r.xr = r.xl + KC * d;
        r.w = r.xr + KC * d;
        return r;
    };
    auto issue = [&](int s, int cls, int64_t chunk) {             // one thread
        RawStage r = stage_ptrs(s);
        const int64_t i0 = chunk * KC;
        fence_proxy_async();
        mbar_expect_tx(&bars[s], raw_bytes);
        bulk_g2s(r.L, L + i0 * chi_l, (uint32_t)(KC * chi_l * sizeof(double)), &bars[s]);
        bulk_g2s(r.R, R + i0 * chi_r, (uint32_t)(KC * chi_r * sizeof(double)), &bars[s]);
        bulk_g2s(r.xl, xl + i0 * d, (uint32_t)(KC * d * sizeof(double)), &bars[s]);
        bulk_g2s(r.xr, xr + i0 * d, (uint32_t)(KC * d * sizeof(double)), &bars[s]);
        bulk_g2s(r.w, w + (int64_t)cls * wstride + i0, (uint32_t)(KC * sizeof(double)), &bars[s]);
    };

    for (int sg = cta_ptr[blockIdx.x]; sg < cta_ptr[blockIdx.x + 1]; sg++) {
        const GradSeg seg = segs[sg];
        const int p0 = seg.tp * TP, q0 = seg.tq * TQ;
        // operand build.  Symmetric tiles (TP == TQ == 128): thread t owns column t & 127 of BOTH tiles and every
        // second row, so all warps carry the same P + Q mix.  Asymmetric tiles: thread t owns one column of the
        // concatenated [P | Q] column space for all KC rows; both kinds run one instruction stream
        // val = f0[i] * f1[i*d + s] * f2[i*chi + link] (Q threads read f0 from a vector of ones).
        constexpr bool SYM = (TP == TQ);
        const int pl = SYM ? (tid & (TP - 1)) : tid, ihalf = tid >> 7;
        const bool is_p = SYM ? true : tid < TP;
        const int ql = SYM ? pl : tid - TP;
        const int pp = p0 + pl, qq = q0 + ql;
        const bool pok = is_p && pp < Dl, qok = (SYM || !is_p) && qq < Dr;
        const int pa = pok ? pp / d : 0, ps = pok ? pp - pa * d : 0;
        const int qb = qok ? qq / d : 0, qs = qok ? qq - qb * d : 0;

        auto build = [&](int s, int buf) {
            RawStage r = stage_ptrs(s);
            double* P = Pt + buf * KC * LDT;
            double* Q = Qt + buf * KC * LDT;
            if (SYM) {
#pragma unroll
                for (int rr = 0; rr < KC / 2; rr++) {
                    const int i = 2 * rr + ihalf;
                    const double wi = r.w[i];
                    P[i * LDT + pl] = pok ? wi * r.xl[i * d + ps] * r.L[i * chi_l + pa] : 0.0;
                    Q[i * LDT + pl] = qok ? r.xr[i * d + qs] * r.R[i * chi_r + qb] : 0.0;
                }
            } else {
                const double* f1 = is_p ? r.xl + ps : r.xr + qs;
                const double* f2 = is_p ? r.L + pa : r.R + qb;
                const int s2 = is_p ? chi_l : chi_r;
                double* dst = is_p ? P + pl : Q + ql;
                const bool ok = is_p ? pok : qok;
#pragma unroll
                for (int i = 0; i < KC; i++) {
                    const double f0 = is_p ? r.w[i] : 1.0;
                    dst[i * LDT] = ok ? f0 * f1[i * d] * f2[i * s2] : 0.0;
                }
            }
        };

        double acc[MI][NI][2];
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        const int64_t cbeg = seg.chunk_begin, nch = seg.chunk_end - seg.chunk_begin;   // in KC-sample chunks
        // prologue: raw(0), raw(1) in flight; build(0)
        if (tid == 0) {
            issue(0, seg.cls, cbeg);
            if (nch > 1) issue(1, seg.cls, cbeg + 1);
        }
        mbar_wait(&bars[0], phase[0]);
        phase[0] ^= 1;
        build(0, 0);
        __syncthreads();                                           // built[0] visible, raw[0] free
        if (tid == 0 && nch > 2) issue(0, seg.cls, cbeg + 2);

        // Warps 0-3 build the next operand tiles and then run their DMMAs, warps 4-7 do it the other way round:
        // every SM sub-partition hosts one warp of each kind, so its FP64 tensor pipe always has DMMAs queued
        // while the other warp is on the load/store and multiply work of the build.
        const bool build_first = (warp & 4) == 0;
        for (int64_t j = 0; j < nch; j++) {
            const int cur = (int)(j & 1);
            auto build_next = [&]() {
                if (j + 1 < nch) {
                    const int s = (int)((j + 1) & 1);
                    mbar_wait(&bars[s], phase[s]);
                    phase[s] ^= 1;
                    build(s, cur ^ 1);
                }
            };
            if (build_first) build_next();
            const double* Pc = Pt + cur * KC * LDT + fc * LDT + wp * (8 * MI) + fr;
            const double* Qc = Qt + cur * KC * LDT + fc * LDT + wq * (8 * NI) + fr;
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; k4++) {
                double a[MI], b[NI];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) a[mi] = Pc[k4 * 4 * LDT + mi * 8];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) b[ni] = Qc[k4 * 4 * LDT + ni * 8];
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
#pragma unroll
                    for (int ni = 0; ni < NI; ni++)
                        dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            if (!build_first) build_next();
            __syncthreads();                                       // built[cur] consumed, raw[(j+1)&1] free
            if (tid == 0 && j + 3 < nch) issue((int)((j + 1) & 1), seg.cls, cbeg + j + 3);
        }

        // partial tile: dense TP x TQ, p fastest
        double* dst = part + (size_t)seg.slot * TP * TQ;
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                const int p = wp * (8 * MI) + mi * 8 + fr;
                const int q = wq * (8 * NI) + ni * 8 + 2 * fc;
                dst[p + TP * q] = acc[mi][ni][0];
                dst[p + TP * (q + 1)] = acc[mi][ni][1];
            }
        __syncthreads();
    }
}

// G[c][p + Dl*q] = sum over the tile's segments, fixed order
__global__ void __launch_bounds__(256)
grad_reduce_kernel(const double* __restrict__ part, const int* __restrict__ tile_slot, int ntp, int ntq,
                   int Dl, int Dr, int TP, int TQ, double* __restrict__ G) {
    const int tile = blockIdx.x;                                   // cls*ntp*ntq + tp*ntq + tq
    const int cls = tile / (ntp * ntq);
    const int rem = tile - cls * ntp * ntq;
    const int tp = rem / ntq, tq = rem - tp * ntq;
    const int s0 = tile_slot[tile], s1 = tile_slot[tile + 1];
    double* Gc = G + (size_t)cls * Dl * Dr;
    for (int e = threadIdx.x; e < TP * TQ; e += blockDim.x) {
        const int pl = e % TP, ql = e / TP;
        const int p = tp * TP + pl, q = tq * TQ + ql;
        if (p >= Dl || q >= Dr) continue;
        double s = 0.0;
        for (int sl = s0; sl < s1; sl++) s += part[(size_t)sl * TP * TQ + e];
        Gc[(size_t)p + (size_t)Dl * q] = s;
    }
}
}  // namespace

// Host side: build the stream-K table for `ncls` class segments.  cls_begin/cls_end are sample
// ranges (KLD: the class's own samples; MSE: all samples for every class).
int launch_bond_grad(mpst_ctx* c, const double* xl, const double* xr, const double* L, const double* R,
                     int d, int chi_l, int chi_r, const int64_t* cls_begin, const int64_t* cls_end,
                     int ncls, double* G) {
    const int Dl = d * chi_l, Dr = d * chi_r;
    // tile edge 96 or 128 per dimension, whichever pads less (ties -> 128)
    // tile shapes: 128 x 128 by default; 160 x 96 when that removes the padding (d*chi = 480 at config B:
    // 3 x 5 exact tiles instead of 4 x 4 tiles of which 12 % is padding)
    int TP = 128, TQ = 128;
    if (!c->flag[F_GRAD_T128]) {
        const double w128 = (double)((Dl + 127) / 128 * 128) * ((Dr + 127) / 128 * 128);
        const double w160 = (double)((Dl + 159) / 160 * 160) * ((Dr + 95) / 96 * 96);
        if (w160 < 0.95 * w128) { TP = 160; TQ = 96; }
    }
    const int ntp = (Dl + TP - 1) / TP, ntq = (Dr + TQ - 1) / TQ;
    const int ntiles = ncls * ntp * ntq;
    const int ncta = c->sm_count;
    // chunks per class
    std::vector<int64_t> cb(ncls), ce(ncls);
    int64_t total = 0;
    // chunk size: 32 samples per pipeline step when the staging buffers fit in shared memory, else 16
    const int LDT = std::max(TP, TQ) + 4;
    const size_t smem32 = 16 + sizeof(double) * (2 * (size_t)32 * (chi_l + chi_r + 2 * d + 1) + 4 * (size_t)32 * LDT);
    const int KC = (smem32 <= 227 * 1024 && c->flag[F_GRAD_KC] == 32) ? 32 : 16;   // measured: 16 is faster (24.5 vs 23.8 TFLOP/s)
    for (int k = 0; k < ncls; k++) {
        cb[k] = cls_begin[k] / KC;
        ce[k] = (cls_end[k] + KC - 1) / KC;
        if (cls_end[k] <= cls_begin[k]) ce[k] = cb[k];
        total += (ce[k] - cb[k]) * ntp * ntq;
    }
    if (total == 0) {
        CUDA_TRY(c, cudaMemsetAsync(G, 0, sizeof(double) * (size_t)ncls * Dl * Dr, c->stream));
        return MPST_OK;
    }
    // The schedule depends only on (tile shape, tile counts, class chunk ranges): built once, kept on the device
    const int phases = c->flag[F_GRAD_PHASES];
    std::vector<int64_t> key = {2, TP, TQ, KC, ncta, ntp, ntq, ncls, phases};
    for (int k = 0; k < ncls; k++) { key.push_back(cb[k]); key.push_back(ce[k]); }
    SegTable* tab = segtable_find(c, key);
    if (!tab) {
        std::vector<GradSeg> hsegs;
        std::vector<int> hcta, hslot;
        std::vector<std::array<int, 3>> units;
        for (int tile = 0; tile < ntiles; tile++) {
            const int rem = tile % (ntp * ntq);
            units.push_back({tile / (ntp * ntq), rem / ntq, rem % ntq});
        }
        build_streamk_table(ncta, phases, units, cb, ce, hsegs, hcta, hslot);
        TRY(segtable_add(c, key, hsegs, hcta, hslot, &tab));
    }
    const int nseg = tab->nseg;
    TRY(ensure_buf(c, &c->part, &c->partcap, (size_t)nseg * TP * TQ));

    const size_t smem = 16 + sizeof(double) * (2 * (size_t)KC * (chi_l + chi_r + 2 * d + 1) + 4 * (size_t)KC * LDT);
    void (*kern)(const double*, const double*, const double*, const double*, const double*, int64_t, int, int, int,
                 const GradSeg*, const int*, double*) = nullptr;
    if (KC == 32) kern = TP == 128 ? bond_grad_kernel<32, 4, 8> : bond_grad_kernel<32, 5, 6>;
    else kern = TP == 128 ? bond_grad_kernel<16, 4, 8> : bond_grad_kernel<16, 5, 6>;
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    c->last[L_GRAD_KERNEL] = 2;
    c->last[L_GRAD_TILE_LAUNCHES]++;
    c->last[L_GRAD_VARIANT] = TP * 1000 + TQ;
    prof_begin(c, MPST_T_GRADK);
    kern<<<ncta, 256, smem, c->stream>>>(xl, xr, L, R, c->w, c->Npad, d, chi_l, chi_r, tab->segs, tab->cta_ptr, c->part);
    prof_end(c, MPST_T_GRADK);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    grad_reduce_kernel<<<ntiles, 256, 0, c->stream>>>(c->part, tab->tile_slot, ntp, ntq, Dl, Dr, TP, TQ, G);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}
