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
    extern __shared__ double sm[];                       // [256][width+1] would conflict-pad; we
    const int width = (BASIS >= MPST_BASIS_FOURIER && BASIS <= MPST_BASIS_SAHAND) ? 2 * d : d;
    const int64_t base = (int64_t)blockIdx.x * blockDim.x;
    const int64_t i = base + threadIdx.x;
    double v[2 * MPST_MAX_D];
    if (i < n) {
        encode_point<BASIS>(x[i], d, v);
        // stage: thread-major rows of `width`, odd pitch to avoid bank conflicts
        double* row = sm + (size_t)threadIdx.x * (width | 1);
        for (int k = 0; k < width; k++) row[k] = v[k];
    }
    __syncthreads();
    const int64_t cnt = min((int64_t)blockDim.x, n - base);
    if (cnt <= 0) return;
    if (ldo == width) {                                  // dense output: coalesced copy-out
        const int64_t total = cnt * width;
        double* dst = out + base * width;
        for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
            int r = (int)(e / width), k = (int)(e - (int64_t)r * width);
            dst[e] = sm[(size_t)r * (width | 1) + k];
        }
    } else {
        for (int64_t e = threadIdx.x; e < cnt * width; e += blockDim.x) {
            int r = (int)(e / width), k = (int)(e - (int64_t)r * width);
            out[(base + r) * ldo + k] = sm[(size_t)r * (width | 1) + k];
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
    const size_t smem = (size_t)threads * (width | 1) * sizeof(double);
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
