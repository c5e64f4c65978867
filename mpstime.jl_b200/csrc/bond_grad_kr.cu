// K2 (gradient half), register-operand variant: the same GEMM as bond_grad.cu
//     G_c[s + d*a, t + d*b] = sum_i w_c[i] xl_i[s] L_i[a] xr_i[t] R_i[b]
// but the Khatri-Rao operands are never built in shared memory.  On sm_100 a DMMA holds its sub-partition's
// dispatch port for its whole 16-cycle issue, so every non-DMMA instruction of any warp on that sub-partition
// delays the tensor pipe; the operand build of bond_grad.cu (2 loads + multiply + store per tile element, plus a
// block barrier per chunk) is most of that overhead.  Here a warp owns a (8 MA links x S sites) x (8 NB links x
// T sites) block of G: its A/B fragments come straight from the raw L / R rows that the bulk-TMA ring delivers,
// and are scaled in registers by the per-sample site values (MA*S + NB*T multiplies for MA*S*NB*T DMMAs per
// k-step of 4 samples; the sample weight rides on the A fragment).  No operand tiles, no block barrier: warp 0
// issues the bulk-TMA copies of the next 64-sample stage (3-stage ring), every warp waits on the stage's
// full-mbarrier and releases it on the empty-mbarrier.
// Schedule: stream-K over (class, group of 8 warp blocks, 16-sample chunk), one contiguous range per CTA
// (grid = #SMs), deterministic segment reduction in a second kernel, exactly as in bond_grad.cu.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "mpst_common.cuh"
#include "dmma.cuh"
#include "streamk.h"

namespace {
// KC = samples per pipeline stage, SR = raw stages: 64 x 3 where the ring fits in shared memory (config B:
// 29.4 vs 27.5 TFLOP/s with 16-sample stages), 32 x 4 for wide links (d = 16, chi = 64: 160 doubles per sample).

struct KrGeom {
    int nab, nsg, nbb, ntg;  // link blocks / site groups per side
    int units;               // nab*nsg*nbb*ntg warp blocks per class
    int ngroups;             // ceil(units / warps per CTA)
};

template <int MA, int S, int NB, int T, int NW, int KC, int SR>
__global__ void __launch_bounds__(32 * NW, 1)
bond_grad_kr_kernel(const double* __restrict__ xl, const double* __restrict__ xr,
                    const double* __restrict__ L, const double* __restrict__ R,
                    const double* __restrict__ w, int64_t wstride, int d, int chi_l, int chi_r, KrGeom geo,
                    const GradSeg* __restrict__ segs, const int* __restrict__ cta_ptr,
                    double* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char smraw[];
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smraw);      // [SR] TMA landed
    uint64_t* raw_empty = raw_full + SR;                          // [SR] all 8 warps done reading
    double* base = reinterpret_cast<double*>(smraw + 128);
    // rows stay contiguous (one bulk copy per operand and stage): copying row by row into a padded, conflict-free
    // pitch was measured slower (35 small copies per stage saturate the copy engine; the 2-way conflicts cost less)
    const int ldl = chi_l, ldr = chi_r;
    constexpr int LOOK = SR - 2;                                  // stages in flight ahead of the compute cursor
    const int stage_sz = KC * (ldl + ldr + 2 * d) + KC;           // L | R | xl | xr | w
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;

    if (tid == 0) {
        for (int s = 0; s < SR; s++) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], NW); }
        mbar_fence_init();
    }
    __syncthreads();
    const int seg_begin = cta_ptr[blockIdx.x], seg_end = cta_ptr[blockIdx.x + 1];

    // ---- TMA issue cursor (warp 0 only): runs LOOK chunks ahead of the compute cursor -----------------------
    const uint32_t bytes = (uint32_t)(sizeof(double) * (KC * (chi_l + chi_r + 2 * d) + KC));
    int iseg = seg_begin, icls = 0, g_issue = 0;
    int64_t ichunk = 0, iend = 0;
    if (warp == 0 && iseg < seg_end) { const GradSeg sg = segs[iseg]; ichunk = sg.chunk_begin; iend = sg.chunk_end; icls = sg.cls; }
    auto issue_one = [&]() {                                       // whole warp 0
        if (iseg >= seg_end) return;
        const int s = g_issue % SR;
        if (lane == 0) {
            mbar_wait(&raw_empty[s], ((g_issue / SR) & 1) ^ 1);
            fence_proxy_async();
            mbar_expect_tx(&raw_full[s], bytes);
        }
        __syncwarp();
        double* sL = base + (size_t)s * stage_sz;
        double* sR = sL + KC * ldl;
        double* sxl = sR + KC * ldr;
        double* sxr = sxl + KC * d;
        double* sw = sxr + KC * d;
        const int64_t i0 = ichunk * KC;
        // one row per lane: rows land with a pitch == 4 (mod 16) doubles -> conflict-free fragment loads
        if (lane == 3) bulk_g2s(sL, L + i0 * chi_l, (uint32_t)(KC * chi_l * sizeof(double)), &raw_full[s]);
        if (lane == 4) bulk_g2s(sR, R + i0 * chi_r, (uint32_t)(KC * chi_r * sizeof(double)), &raw_full[s]);
        if (lane == 0) bulk_g2s(sxl, xl + i0 * d, (uint32_t)(KC * d * sizeof(double)), &raw_full[s]);
        if (lane == 1) bulk_g2s(sxr, xr + i0 * d, (uint32_t)(KC * d * sizeof(double)), &raw_full[s]);
        if (lane == 2) bulk_g2s(sw, w + (int64_t)icls * wstride + i0, (uint32_t)(KC * sizeof(double)), &raw_full[s]);
        g_issue++;
        if (++ichunk >= iend) {
            if (++iseg < seg_end) { const GradSeg sg = segs[iseg]; ichunk = sg.chunk_begin; iend = sg.chunk_end; icls = sg.cls; }
        }
    };
    if (warp == 0)
        for (int k = 0; k < LOOK; k++) issue_one();

    int g = 0;
    for (int sgi = seg_begin; sgi < seg_end; sgi++) {
        const GradSeg seg = segs[sgi];
        const int64_t nch = seg.chunk_end - seg.chunk_begin;
        const int unit = seg.tp * NW + warp;
        const bool live = unit < geo.units;
        // unit -> (a block, s group, b block, t group)
        int u = live ? unit : 0;
        const int tg = u % geo.ntg; u /= geo.ntg;
        const int sg_ = u % geo.nsg; u /= geo.nsg;
        const int bb = u % geo.nbb; u /= geo.nbb;
        const int ab = u;
        const int a0 = ab * (8 * MA) + fr, b0 = bb * (8 * NB) + fr;
        int aoff[MA], boff[NB];
        bool aok[MA], bok[NB];
#pragma unroll
        for (int ma = 0; ma < MA; ma++) { aok[ma] = a0 + 8 * ma < chi_l; aoff[ma] = aok[ma] ? a0 + 8 * ma : 0; }
#pragma unroll
        for (int nb = 0; nb < NB; nb++) { bok[nb] = b0 + 8 * nb < chi_r; boff[nb] = bok[nb] ? b0 + 8 * nb : 0; }
        const int s0 = sg_ * S, t0 = tg * T;

        double acc[MA * S][NB * T][2];
#pragma unroll
        for (int x = 0; x < MA * S; x++)
#pragma unroll
            for (int y = 0; y < NB * T; y++) acc[x][y][0] = acc[x][y][1] = 0.0;

        for (int64_t j = 0; j < nch; j++, g++) {
            if (warp == 0) issue_one();                            // chunk g + LOOK
            const int st = g % SR;
            mbar_wait(&raw_full[st], (g / SR) & 1);
            if (live) {
                const double* sL = base + (size_t)st * stage_sz;
                const double* sR = sL + KC * ldl;
                const double* sxl = sR + KC * ldr;
                const double* sxr = sxl + KC * d;
                const double* sw = sxr + KC * d;
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; k4++) {
                    const int i = 4 * k4 + fc;                     // this lane's sample of the k-step
                    double af[MA], bf[NB], cl[S], cr[T];
                    const double wi = sw[i];
#pragma unroll
                    for (int ma = 0; ma < MA; ma++) af[ma] = sL[i * ldl + aoff[ma]];
#pragma unroll
                    for (int nb = 0; nb < NB; nb++) bf[nb] = sR[i * ldr + boff[nb]];
#pragma unroll
                    for (int s = 0; s < S; s++) cl[s] = sxl[i * d + s0 + s];
#pragma unroll
                    for (int t = 0; t < T; t++) cr[t] = sxr[i * d + t0 + t];
                    // the sample weight rides on the A fragment: MA multiplies instead of S
#pragma unroll
                    for (int ma = 0; ma < MA; ma++) af[ma] = aok[ma] ? af[ma] * wi : 0.0;
#pragma unroll
                    for (int nb = 0; nb < NB; nb++) bf[nb] = bok[nb] ? bf[nb] : 0.0;
                    double as[MA * S], bs[NB * T];
#pragma unroll
                    for (int ma = 0; ma < MA; ma++)
#pragma unroll
                        for (int s = 0; s < S; s++) as[ma * S + s] = af[ma] * cl[s];
#pragma unroll
                    for (int nb = 0; nb < NB; nb++)
#pragma unroll
                        for (int t = 0; t < T; t++) bs[nb * T + t] = bf[nb] * cr[t];
#pragma unroll
                    for (int x = 0; x < MA * S; x++)
#pragma unroll
                        for (int y = 0; y < NB * T; y++) dmma_8x8x4(acc[x][y][0], acc[x][y][1], as[x], bs[y]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[st]);
        }
        if (live) {
            // partial block of this unit: local row (ma*S + s)*8 + a_l, local column (nb*T + t)*8 + b_l
            constexpr int RU = MA * S * 8, CU = NB * T * 8;
            double* dst = part + ((size_t)seg.slot * NW + warp) * RU * CU;
#pragma unroll
            for (int x = 0; x < MA * S; x++)
#pragma unroll
                for (int y = 0; y < NB * T; y++) {
                    const int lr = x * 8 + fr, lc = y * 8 + 2 * fc;
                    dst[lr + RU * lc] = acc[x][y][0];
                    dst[lr + RU * (lc + 1)] = acc[x][y][1];
                }
        }
    }
}

// G[c][p + Dl*q] = sum over the group's segments (fixed order), scattered from the unit-local layout
template <int MA, int S, int NB, int T, int NW>
__global__ void __launch_bounds__(256)
grad_kr_reduce_kernel(const double* __restrict__ part, const int* __restrict__ grp_slot, KrGeom geo, int d, int chi_l,
                      int chi_r, double* __restrict__ G) {
    constexpr int RU = MA * S * 8, CU = NB * T * 8;
    const int grp = blockIdx.x;                                    // cls * ngroups + group
    const int cls = grp / geo.ngroups, group = grp - cls * geo.ngroups;
    const int s0 = grp_slot[grp], s1 = grp_slot[grp + 1];
    const int Dl = d * chi_l, Dr = d * chi_r;
    double* Gc = G + (size_t)cls * Dl * Dr;
    {
        const int wu = blockIdx.y;                                 // one block per warp block of the group
        const int unit = group * NW + wu;
        if (unit >= geo.units) return;
        int u = unit;
        const int tg = u % geo.ntg; u /= geo.ntg;
        const int sg_ = u % geo.nsg; u /= geo.nsg;
        const int bb = u % geo.nbb; u /= geo.nbb;
        const int ab = u;
        for (int e = threadIdx.x; e < RU * CU; e += blockDim.x) {
            const int lr = e % RU, lc = e / RU;
            const int x = lr >> 3, al = lr & 7, y = lc >> 3, bl = lc & 7;
            const int a = (ab * MA + x / S) * 8 + al, s = sg_ * S + x % S;
            const int b = (bb * NB + y / T) * 8 + bl, t = tg * T + y % T;
            if (a >= chi_l || b >= chi_r || s >= d || t >= d) continue;
            double sum = 0.0;
            for (int sl = s0; sl < s1; sl++) sum += part[((size_t)sl * NW + wu) * RU * CU + e];
            Gc[(size_t)(s + d * a) + (size_t)Dl * (t + d * b)] = sum;
        }
    }
}

template <int MA, int S, int NB, int T, int NW, int KC, int SR>
int launch_kr(mpst_ctx* c, const double* xl, const double* xr, const double* L, const double* R, int d, int chi_l,
              int chi_r, const int64_t* cls_begin, const int64_t* cls_end, int ncls, double* G, size_t smem) {
    KrGeom geo;
    geo.nab = (chi_l + 8 * MA - 1) / (8 * MA);
    geo.nbb = (chi_r + 8 * NB - 1) / (8 * NB);
    geo.nsg = (d + S - 1) / S;
    geo.ntg = (d + T - 1) / T;
    geo.units = geo.nab * geo.nsg * geo.nbb * geo.ntg;
    geo.ngroups = (geo.units + NW - 1) / NW;
    const int ncta = c->sm_count;
    const int ngrp_total = ncls * geo.ngroups;
    std::vector<int64_t> cb(ncls), ce(ncls);
    int64_t total = 0;
    for (int k = 0; k < ncls; k++) {
        cb[k] = cls_begin[k] / KC;
        ce[k] = (cls_end[k] + KC - 1) / KC;
        if (cls_end[k] <= cls_begin[k]) ce[k] = cb[k];
        total += (ce[k] - cb[k]) * geo.ngroups;
    }
    if (total == 0) {
        CUDA_TRY(c, cudaMemsetAsync(G, 0, sizeof(double) * (size_t)ncls * d * chi_l * d * chi_r, c->stream));
        return MPST_OK;
    }
    // The schedule depends only on (kernel variant, group count, class chunk ranges): built once, kept on the device
    const int phases = c->flag[F_GRAD_PHASES];
    std::vector<int64_t> key = {1, MA, S, NB, T, NW, KC, SR, ncta, geo.ngroups, ncls, phases};
    for (int k = 0; k < ncls; k++) { key.push_back(cb[k]); key.push_back(ce[k]); }
    SegTable* tab = segtable_find(c, key);
    if (!tab) {
        std::vector<GradSeg> hsegs;
        std::vector<int> hcta, hslot;
        std::vector<std::array<int, 3>> units;
        for (int grp = 0; grp < ngrp_total; grp++) units.push_back({grp / geo.ngroups, grp % geo.ngroups, 0});
        build_streamk_table(ncta, phases, units, cb, ce, hsegs, hcta, hslot);
        TRY(segtable_add(c, key, hsegs, hcta, hslot, &tab));
    }
    const int nseg = tab->nseg;
    constexpr int RU = MA * S * 8, CU = NB * T * 8;
    TRY(ensure_buf(c, &c->part, &c->partcap, (size_t)nseg * NW * RU * CU));
    c->last[L_GRAD_KERNEL] = 1;
    c->last[L_GRAD_KR_LAUNCHES]++;
    c->last[L_GRAD_VARIANT] = MA * 10000 + S * 1000 + KC * 10 + SR;
    auto kern = bond_grad_kr_kernel<MA, S, NB, T, NW, KC, SR>;
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(c, MPST_T_GRADK);
    kern<<<ncta, 32 * NW, smem, c->stream>>>(xl, xr, L, R, c->w, c->Npad, d, chi_l, chi_r, geo, tab->segs, tab->cta_ptr, c->part);
    prof_end(c, MPST_T_GRADK);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    grad_kr_reduce_kernel<MA, S, NB, T, NW><<<dim3(ngrp_total, NW), 256, 0, c->stream>>>(c->part, tab->tile_slot, geo, d, chi_l, chi_r, G);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}
}  // namespace

// Returns MPST_OK and sets *handled when the register-operand kernel covers this shape; otherwise the caller
// uses the shared-memory-tile kernel of bond_grad.cu.
int launch_bond_grad_kr(mpst_ctx* c, const double* xl, const double* xr, const double* L, const double* R, int d,
                        int chi_l, int chi_r, const int64_t* cls_begin, const int64_t* cls_end, int ncls, double* G,
                        bool* handled) {
    *handled = false;
    if (c->flag[F_GRAD_NOKR]) return MPST_OK;
    if ((chi_l & 1) || (chi_r & 1) || chi_l < 8 || chi_r < 8) return MPST_OK;       // 16-byte rows for the bulk copies
    auto ring = [&](int kc, int sr) { return 128 + sizeof(double) * (size_t)sr * (kc * (chi_l + chi_r + 2 * d) + kc); };
    const int forced = c->flag[F_GRAD_KC];
    // stage size x ring depth by what fits: 64 x 4 or 64 x 3 (config B), 32 x 4 (d = 16, chi = 64), 16 x 4 (chi = 128);
    // GRAD_KC = kc * 10 + sr forces a variant (A/B measurements)
    const int lim = 227 * 1024;
    int kc, sr;
    if (forced >= 100) { kc = forced / 10; sr = forced % 10; }
    else if (ring(64, 4) <= (size_t)lim && c->flag[F_GRAD_KC] == 0 && c->kr_deep) { kc = 64; sr = 4; }
    else if (ring(64, 3) <= (size_t)lim && forced != 32 && forced != 16) { kc = 64; sr = 3; }
    else if (ring(32, 4) <= (size_t)lim && forced != 16) { kc = 32; sr = 4; }
    else { kc = 16; sr = 4; }
    const size_t smem = ring(kc, sr);
    if (smem > (size_t)lim) return MPST_OK;
#define KR_ARGS c, xl, xr, L, R, d, chi_l, chi_r, cls_begin, cls_end, ncls, G, smem
#define KR_DISPATCH(MA_, S_, NB_, T_)                                                                   \
    do {                                                                                                \
        *handled = true;                                                                                \
        if (kc == 64 && sr == 4) return launch_kr<MA_, S_, NB_, T_, 8, 64, 4>(KR_ARGS);                 \
        if (kc == 64) return launch_kr<MA_, S_, NB_, T_, 8, 64, 3>(KR_ARGS);                            \
        if (kc == 32 && sr == 6) return launch_kr<MA_, S_, NB_, T_, 8, 32, 6>(KR_ARGS);                 \
        if (kc == 32) return launch_kr<MA_, S_, NB_, T_, 8, 32, 4>(KR_ARGS);                            \
        return launch_kr<MA_, S_, NB_, T_, 8, 16, 4>(KR_ARGS);                                          \
    } while (0)
    if (d % 6 == 0) KR_DISPATCH(1, 6, 1, 6);
    if (d % 4 == 0) KR_DISPATCH(2, 4, 1, 4);
#undef KR_DISPATCH
#undef KR_ARGS
    return MPST_OK;
}
