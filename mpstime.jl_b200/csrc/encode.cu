// K1: basis encoding of scalar time points (replaces encode_TS -> Basis.encode,
// reference Encodings/encodings.jl:18-27 and Encodings/bases.jl:13-92).
// HBM-bound elementwise kernel: 8 B in, 8*d B out per point (16*d for the complex bases).
// One thread per point runs the recurrence in registers; a warp then writes its 32*d outputs
// through shared memory so that global stores are fully coalesced 16-byte vectors.
#include "mpst_common.cuh"
#include "encode_device.cuh"

template <int BASIS>
__global__ void __launch_bounds__(256) encode_kernel(const double* __restrict__ x, int64_t n, int d,
                                                     double* __restrict__ out, int64_t ldo) {
    extern __shared__ double sm[];                       // [256][width|1] staging rows, then 3*32 recurrence constants
    constexpr bool LEG = BASIS == MPST_BASIS_LEGENDRE_NO_NORM || BASIS == MPST_BASIS_LEGENDRE_NORM;
    const int width = (BASIS >= MPST_BASIS_FOURIER && BASIS <= MPST_BASIS_SAHAND) ? 2 * d : d;
    const int pitch = width | 1;                         // odd pitch: conflict-free row writes
    const int64_t base = (int64_t)blockIdx.x * blockDim.x;
    const int64_t i = base + threadIdx.x;
    double* row = sm + (size_t)threadIdx.x * pitch;
    if (LEG) {
        // Bonnet recurrence P_{l+1} = a_l x P_l - b_l P_{l-1} with a_l = (2l+1)/(l+1), b_l = l/(l+1) and the
        // normalisation sqrt((2l+1)/2) [times the Legendre_Norm factor, bases.jl:86-89]: the divisions and square
        // roots are done once per block (they were most of the instructions per point), every value goes straight
        // to its staging slot (no per-thread array, no local memory)
        double* ca = sm + (size_t)blockDim.x * pitch;
        double* cb = ca + MPST_MAX_D;
        double* cn = cb + MPST_MAX_D;
        if (threadIdx.x < d) {
            const int l = threadIdx.x;
            ca[l] = (double)(2 * l + 1) / (double)(l + 1);
            cb[l] = (double)l / (double)(l + 1);
            double nrm = sqrt((double)(2 * l + 1) * 0.5);
            if (BASIS == MPST_BASIS_LEGENDRE_NORM) nrm /= sqrt(sqrt((double)(2 * d + 1) * 0.5) * (double)d);
            cn[l] = nrm;
        }
        __syncthreads();
        if (i < n) {
            const double xv = x[i];
            double pm = 0.0, p = 1.0;
            row[0] = cn[0];
            for (int l = 0; l + 1 < d; l++) {
                const double pn = ca[l] * xv * p - cb[l] * pm;
                pm = p;
                p = pn;
                row[l + 1] = cn[l + 1] * p;
            }
        }
    } else if (i < n) {
        double v[2 * MPST_MAX_D];
        encode_point<BASIS>(x[i], d, v);
        for (int k = 0; k < width; k++) row[k] = v[k];
    }
    __syncthreads();
    const int64_t cnt = min((int64_t)blockDim.x, n - base);
    if (cnt <= 0) return;
    if (ldo == width) {                                  // dense output: coalesced copy-out, 128-bit stores when aligned
        const int64_t total = cnt * width;
        double* dst = out + base * width;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            const int64_t pairs = total >> 1;
            for (int64_t e2 = threadIdx.x; e2 < pairs; e2 += blockDim.x) {
                const int64_t e = 2 * e2;
                const int r0 = (int)(e / width), k0 = (int)(e - (int64_t)r0 * width);
                const int r1 = (k0 + 1 == width) ? r0 + 1 : r0, k1 = (k0 + 1 == width) ? 0 : k0 + 1;
                double2 v2;
                v2.x = sm[(size_t)r0 * pitch + k0];
                v2.y = sm[(size_t)r1 * pitch + k1];
                reinterpret_cast<double2*>(dst)[e2] = v2;
            }
            if ((total & 1) && threadIdx.x == 0) {
                const int64_t e = total - 1;
                dst[e] = sm[(size_t)(e / width) * pitch + (e % width)];
            }
        } else {
            for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
                int r = (int)(e / width), k = (int)(e - (int64_t)r * width);
                dst[e] = sm[(size_t)r * pitch + k];
            }
        }
    } else {
        for (int64_t e = threadIdx.x; e < cnt * width; e += blockDim.x) {
            int r = (int)(e / width), k = (int)(e - (int64_t)r * width);
            out[(base + r) * ldo + k] = sm[(size_t)r * pitch + k];
        }
    }
}

int launch_encode(mpst_ctx* c, int basis, int d, const double* x, int64_t n, double* out, int64_t ldo) {
    if (n <= 0) return MPST_OK;
    if (d < 1 || d > MPST_MAX_D) { c->err = "encode: d out of range"; return MPST_E_INVALID; }
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    const bool cplx = basis >= MPST_BASIS_FOURIER && basis <= MPST_BASIS_SAHAND;
    const int width = cplx ? 2 * d : d;
    const size_t smem = ((size_t)threads * (width | 1) + 3 * MPST_MAX_D) * sizeof(double);
#define ENC_CASE(BID)                                                                            \
    case BID:                                                                                    \
        cudaFuncSetAttribute(encode_kernel<BID>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                             (int)smem);                                                         \
        encode_kernel<BID><<<(unsigned)blocks, threads, smem, c->stream>>>(x, n, d, out, ldo);   \
        break;
    switch (basis) {
        ENC_CASE(MPST_BASIS_LEGENDRE_NO_NORM)
        ENC_CASE(MPST_BASIS_LEGENDRE_NORM)
        ENC_CASE(MPST_BASIS_FOURIER)
        ENC_CASE(MPST_BASIS_STOUDENMIRE)
        ENC_CASE(MPST_BASIS_SAHAND)
        ENC_CASE(MPST_BASIS_UNIFORM)
        default:
            c->err = "encode: unknown basis id";
            return MPST_E_INVALID;
    }
#undef ENC_CASE
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}
