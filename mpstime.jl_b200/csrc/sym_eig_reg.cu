// K5, Rayleigh-Ritz eigen-solver with the matrix resident in REGISTERS.
//
// Same mathematics as sym_eig_kernel<.., LMODE = true> (svd_subspace.cu): one-sided (Hestenes) Jacobi on the columns
// of the lower Cholesky factor L of the Ritz matrix H = L L^T; right rotations J make the columns of L J mutually
// orthogonal, L J = W Sigma, so w_i = (L J)_i / ||(L J)_i|| and lambda_i = ||(L J)_i||^2.  Same rotation formula, same
// rotation floor, same stopping rule.  What changes is where the columns live.  The shared-memory version moves
// every column through shared memory once per round (200 KB per round at p = 112: 1 570 cycles of shared-memory
// bandwidth, ncu: 3 050 cycles per round).  Here a half-warp owns 8 adjacent column SLOTS for the whole solve, each
// lane holding rows l, l+16, ... of all 8 (56 doubles at p = 112), and the pairs come from the odd-even
// transposition ordering:
//     phase A: slots (0,1) (2,3) (4,5) (6,7) of every half-warp           -- entirely lane-local
//     phase B: slots (1,2) (3,4) (5,6) and the pair that straddles two half-warps (7 | 0)
// with the two columns of a pair SWAPPING slots after their rotation, so that after p phases every pair of columns
// has met exactly once (the transposition network reverses the column order).  Only the straddling column of phase
// B travels through shared memory (1/8 of the matrix every second phase).  Per phase a half-warp reduces its four
// dot products with one transposed butterfly (5 shuffles for 4 values), four lanes compute the four rotations at
// once, and the result (c, s, updated squared norms) comes back through a 128-byte shared-memory broadcast.
// The kernel then sits on the single-SM FP64 rate: 5 flop-instructions per row and pair.
#include "mpst_common.cuh"

namespace {

// 4 per-lane partial sums -> full sums over the half-warp; the lane ends up with the total of value
// idx = 2*bit3 + bit2 of its half-warp lane number
__device__ __forceinline__ double reduce4_hw(const double v[4], int l16) {
    const bool b3 = l16 & 8, b2 = l16 & 4;
    const double k0 = (b3 ? v[2] : v[0]) + __shfl_xor_sync(0xffffffffu, b3 ? v[0] : v[2], 8);
    const double k1 = (b3 ? v[3] : v[1]) + __shfl_xor_sync(0xffffffffu, b3 ? v[1] : v[3], 8);
    double t = (b2 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, b2 ? k0 : k1, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}

// rotation of the pair (alpha = |a|^2, beta = |b|^2, gamma = a.b): out = (c, s, |a'|^2, |b'|^2) with
// a' = c a - s b, b' = s a + c b orthogonal; identity when gamma is at the rotation floor.
__device__ __forceinline__ void jacobi_params(double al, double be, double ga, double thr2, double out[4], bool& big) {
    const double g2 = ga * ga, ab = al * be;
    double c = 1.0, sn = 0.0, na = al, nb = be;
    if (g2 > thr2 * fmax(al, be) && g2 > 1e-30 * ab) {
        // cos/sin of the double angle, then half-angle: two rsqrt, no division
        const double zeta = be - al, beta2 = 2.0 * ga;
        const double inv_r = rsqrt(zeta * zeta + beta2 * beta2);
        const double cos2 = fabs(zeta) * inv_r, sin2 = (zeta >= 0.0 ? beta2 : -beta2) * inv_r;
        const double c2 = 0.5 + 0.5 * cos2;
        const double inv_c = rsqrt(c2);
        c = c2 * inv_c;
        sn = 0.5 * sin2 * inv_c;
        const double cs2 = 2.0 * c * sn * ga, cc = c * c, ss = sn * sn;
        na = cc * al - cs2 + ss * be;
        nb = ss * al + cs2 + cc * be;
        big |= g2 > 1e-16 * ab;                        // cosine between the columns above 1e-8
    }
    out[0] = c; out[1] = sn; out[2] = na; out[3] = nb;
}

// Lm: lower Cholesky factor, column-major P x P (upper triangle ignored), P = 16 NR.  W: normalised columns of L J,
// ev: their squared norms (unsorted).  status[0] |= 2 when 60 sweeps did not converge, status[1] = sweeps used.
template <int NR>
__global__ void __launch_bounds__(32 * NR, 1)
sym_eig_reg_kernel(const double* __restrict__ Lm, double* __restrict__ W, double* __restrict__ ev,
                   int* __restrict__ status) {
    constexpr int P = 16 * NR, NG = 2 * NR;            // order, half-warp groups
    __shared__ __align__(16) double xb[NG][P + 2];     // straddling column of phase B (+ its squared norm at [P])
    __shared__ __align__(16) double prm[NG][4][4];     // (c, s, |a'|^2, |b'|^2) of the group's four pairs
    __shared__ double nbuf[NG][8];
    __shared__ double trp[NG];
    const int tid = threadIdx.x, l16 = tid & 15, g = tid >> 4;
    const int idx = ((l16 & 8) ? 2 : 0) + ((l16 & 4) ? 1 : 0);     // the pair this lane computes the rotation of
    double x[8][NR], nrm[8];
#pragma unroll
    for (int c = 0; c < 8; c++)
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const int i = l16 + 16 * r, j = 8 * g + c;
            x[c][r] = (i >= j) ? Lm[i + (size_t)P * j] : 0.0;
        }

    // squared norms of the group's 8 columns from the data (all lanes get all 8)
    auto refresh_norms = [&]() {
        double v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < NR; r++) s = fma(x[c][r], x[c][r], s);
            v[c] = s;
        }
        const double t0 = reduce4_hw(v, l16), t1 = reduce4_hw(v + 4, l16);
        if ((l16 & 3) == 0) { nbuf[g][idx] = t0; nbuf[g][4 + idx] = t1; }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; c++) nrm[c] = nbuf[g][c];
        __syncwarp();
    };
    refresh_norms();
    if (l16 == 0) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 8; c++) s += nrm[c];
        trp[g] = s;
    }
    __syncthreads();
    double tr = 0.0;
    for (int q = 0; q < NG; q++) tr += trp[q];
    const double thr2 = (2.2e-16 * 2.2e-16) * tr;      // columns of L J carry absolute noise eps*sqrt(trace)

    // one phase on four pairs: lo[k] / hi[k] are the two columns, al / be their squared norms for this lane's pair.
    // Returns through prm[g][*]; the caller applies the rotations (register indices must stay compile-time).
    bool big = false;
    auto solve4 = [&](const double gam[4], double al, double be) {
        const double ga = reduce4_hw(gam, l16);
        double o[4];
        jacobi_params(al, be, ga, thr2, o, big);
        if ((l16 & 3) == 0) {
            *reinterpret_cast<double2*>(&prm[g][idx][0]) = make_double2(o[0], o[1]);
            *reinterpret_cast<double2*>(&prm[g][idx][2]) = make_double2(o[2], o[3]);
        }
        __syncwarp();
    };

#ifdef MPST_KDEBUG
    long long tk[6] = {0, 0, 0, 0, 0, 0}, tq = 0;
#define TK0() tq = clock64()
#define TK(i) do { const long long t_ = clock64(); tk[i] += t_ - tq; tq = t_; } while (0)
#else
#define TK0()
#define TK(i)
#endif
    int sweep = 0;
    for (; sweep < 60; sweep++) {
        if (sweep > 0) refresh_norms();
        big = false;
#pragma unroll 1
        for (int ph = 0; ph < P / 2; ph++) {
            // ---- phase A: (0,1) (2,3) (4,5) (6,7) ----
            {
                TK0();
                double gam[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < NR; r++) s = fma(x[2 * k][r], x[2 * k + 1][r], s);
                    gam[k] = s;
                }
                const double al = idx == 0 ? nrm[0] : idx == 1 ? nrm[2] : idx == 2 ? nrm[4] : nrm[6];
                const double be = idx == 0 ? nrm[1] : idx == 1 ? nrm[3] : idx == 2 ? nrm[5] : nrm[7];
                TK(0);
                solve4(gam, al, be);
                TK(1);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double2 cs = *reinterpret_cast<const double2*>(&prm[g][k][0]);
                    const double2 nn = *reinterpret_cast<const double2*>(&prm[g][k][2]);
#pragma unroll
                    for (int r = 0; r < NR; r++) {
                        const double a = x[2 * k][r], b = x[2 * k + 1][r];
                        x[2 * k][r] = fma(cs.y, a, cs.x * b);            // b' takes the low slot (swap)
                        x[2 * k + 1][r] = fma(-cs.y, b, cs.x * a);       // a'
                    }
                    nrm[2 * k] = nn.y;
                    nrm[2 * k + 1] = nn.x;
                }
                __syncwarp();                                            // prm is rewritten in phase B
                TK(2);
            }
            // ---- phase B: (7 of the left neighbour | 0) (1,2) (3,4) (5,6) ----
            {
                if (g < NG - 1) {
#pragma unroll
                    for (int r = 0; r < NR; r++) xb[g][l16 + 16 * r] = x[7][r];
                    if (l16 == 0) xb[g][P] = nrm[7];
                }
                __syncthreads();
                TK(3);
                double e[NR], ne = 0.0;
#pragma unroll
                for (int r = 0; r < NR; r++) e[r] = g > 0 ? xb[g > 0 ? g - 1 : 0][l16 + 16 * r] : 0.0;
                if (g > 0) ne = xb[g - 1][P];
                double gam[4];
                {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < NR; r++) s = fma(e[r], x[0][r], s);
                    gam[0] = s;
                }
#pragma unroll
                for (int k = 1; k < 4; k++) {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < NR; r++) s = fma(x[2 * k - 1][r], x[2 * k][r], s);
                    gam[k] = s;
                }
                const double al = idx == 0 ? ne : idx == 1 ? nrm[1] : idx == 2 ? nrm[3] : nrm[5];
                const double be = idx == 0 ? nrm[0] : idx == 1 ? nrm[2] : idx == 2 ? nrm[4] : nrm[6];
                solve4(gam, al, be);
                if (g > 0) {
                    const double2 cs = *reinterpret_cast<const double2*>(&prm[g][0][0]);
                    const double2 nn = *reinterpret_cast<const double2*>(&prm[g][0][2]);
#pragma unroll
                    for (int r = 0; r < NR; r++) {
                        const double a = e[r], b = x[0][r];
                        xb[g - 1][l16 + 16 * r] = fma(cs.y, a, cs.x * b);   // b' goes to the neighbour's slot 7
                        x[0][r] = fma(-cs.y, b, cs.x * a);                  // a' stays here
                    }
                    if (l16 == 0) xb[g - 1][P] = nn.y;
                    nrm[0] = nn.x;
                }
#pragma unroll
                for (int k = 1; k < 4; k++) {
                    const double2 cs = *reinterpret_cast<const double2*>(&prm[g][k][0]);
                    const double2 nn = *reinterpret_cast<const double2*>(&prm[g][k][2]);
#pragma unroll
                    for (int r = 0; r < NR; r++) {
                        const double a = x[2 * k - 1][r], b = x[2 * k][r];
                        x[2 * k - 1][r] = fma(cs.y, a, cs.x * b);
                        x[2 * k][r] = fma(-cs.y, b, cs.x * a);
                    }
                    nrm[2 * k - 1] = nn.y;
                    nrm[2 * k] = nn.x;
                }
                TK(4);
                __syncthreads();
                if (g < NG - 1) {
#pragma unroll
                    for (int r = 0; r < NR; r++) x[7][r] = xb[g][l16 + 16 * r];
                    nrm[7] = xb[g][P];
                }
                TK(5);
            }
        }
        // quadratic convergence: a sweep whose largest rotated cosine was <= 1e-8 leaves cosines at ~1e-16
        if (!__syncthreads_or(big ? 1 : 0)) { sweep++; break; }
    }
#ifdef MPST_KDEBUG
    if (tid == 0 || tid == 32 * (NR - 1))
        printf("[eig_reg p=%d tid=%d sweeps=%d] A: dots %lld  reduce+params+bcast %lld  rotate %lld | B: publish+sync %lld  compute %lld  sync+readback %lld (cycles)\n",
               P, tid, sweep, tk[0], tk[1], tk[2], tk[3], tk[4], tk[5]);
#endif
    if (tid == 0) {
        if (sweep >= 60) atomicOr(status, 2);
        status[1] = sweep;
    }
    refresh_norms();
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int j = 8 * g + c;
        const double nn = nrm[c];
        const double inv = nn > 0.0 ? rsqrt(nn) : 0.0;
#pragma unroll
        for (int r = 0; r < NR; r++) W[(size_t)j * P + l16 + 16 * r] = x[c][r] * inv;
        if (l16 == 0) ev[j] = nn;
    }
}

}  // namespace

// p must be a multiple of 16 in [32, 128]; returns false when the shape is not covered (caller uses sym_eig_kernel)
bool launch_sym_eig_reg(int p, const double* L, double* W, double* ev, int* status, cudaStream_t st) {
    switch (p) {
        case 32: sym_eig_reg_kernel<2><<<1, 64, 0, st>>>(L, W, ev, status); return true;
        case 48: sym_eig_reg_kernel<3><<<1, 96, 0, st>>>(L, W, ev, status); return true;
        case 64: sym_eig_reg_kernel<4><<<1, 128, 0, st>>>(L, W, ev, status); return true;
        case 80: sym_eig_reg_kernel<5><<<1, 160, 0, st>>>(L, W, ev, status); return true;
        case 96: sym_eig_reg_kernel<6><<<1, 192, 0, st>>>(L, W, ev, status); return true;
        case 112: sym_eig_reg_kernel<7><<<1, 224, 0, st>>>(L, W, ev, status); return true;
        case 128: sym_eig_reg_kernel<8><<<1, 256, 0, st>>>(L, W, ev, status); return true;
        default: return false;
    }
}
