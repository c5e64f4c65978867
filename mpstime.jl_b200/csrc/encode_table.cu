// K1 for the data-driven / time-dependent real encodings of the reference: the basis functions depend on
// coefficient tables that the host computes ONCE from the training set (kernel density estimates, series projections,
// histogram bins -- Encodings/bases.jl:294-397, splitbases.jl:13-110 stay on the host, as in the reference) and that the
// device evaluates per time point and per site:
//   MPST_BASIS_TABLE_LEGENDRE_PROJ   legendre_encode(x, nds, ds[ti]; norm)       bases.jl:95-108
//   MPST_BASIS_TABLE_SAHAND_LEGENDRE sahand_legendre_encode(x, d, kde, minx, scale, cVecs)   bases.jl:111-129
//   MPST_BASIS_TABLE_SPLIT           project_onto_bins(x, aux_dim, aux_encoder, bins)          splitbases.jl:113-163
// Same shape as encode.cu: one thread per point, d outputs staged through shared memory, coalesced stores.
#include "mpst_common.cuh"
#include "encode_device.cuh"
#include "encode_table_device.cuh"

namespace {

template <int KIND>
__global__ void __launch_bounds__(256) encode_table_kernel(const double* __restrict__ x, int64_t n, int d,
                                                           const int* __restrict__ ip, const double* __restrict__ dp,
                                                           double* __restrict__ out, int64_t ldo) {
    extern __shared__ double sm[];
    const int64_t base = (int64_t)blockIdx.x * blockDim.x;
    const int64_t i = base + threadIdx.x;
    const int pitch = d | 1;
    if (i < n) {
        double v[MPST_MAX_D];
        if (KIND == MPST_BASIS_TABLE_LEGENDRE_PROJ) legendre_proj_point(x[i], d, ip, dp, v);
        else if (KIND == MPST_BASIS_TABLE_SAHAND_LEGENDRE) sahand_legendre_point(x[i], d, ip, dp, v);
        else split_point(x[i], d, ip, dp, v);
        double* row = sm + (size_t)threadIdx.x * pitch;
        for (int k = 0; k < d; k++) row[k] = v[k];
    }
    __syncthreads();
    const int64_t cnt = min((int64_t)blockDim.x, n - base);
    if (cnt <= 0) return;
    for (int64_t e = threadIdx.x; e < cnt * d; e += blockDim.x) {
        const int r = (int)(e / d), k = (int)(e - (int64_t)r * d);
        out[(base + r) * ldo + k] = sm[(size_t)r * pitch + k];
    }
}
}  // namespace

// encoded vectors of `n` points that all belong to time point `site` (0-based)
int launch_encode_site(mpst_ctx* c, int site, const double* x, int64_t n, double* out, int64_t ldo) {
    if (c->basis < MPST_BASIS_TABLE_LEGENDRE_PROJ || c->basis > MPST_BASIS_TABLE_SPLIT) return launch_encode(c, c->basis, c->d, x, n, out, ldo);
    if (n <= 0) return MPST_OK;
    const EncTable& t = c->enc;
    if (t.kind != c->basis || t.d != c->d || !t.ip || !t.dp) { c->err = "encode: no coefficient table set for this encoding (mpst_set_encoding_table)"; return MPST_E_INVALID; }
    if (t.nsites != 1 && (site < 0 || site >= t.nsites)) { c->err = "encode: site outside the coefficient table"; return MPST_E_INVALID; }
    const int s = t.nsites == 1 ? 0 : site;
    const int* ip = t.ip + (size_t)s * t.istride;
    const double* dp = t.dp + (size_t)s * t.dstride;
    const int d = c->d, threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    const size_t smem = (size_t)threads * (d | 1) * sizeof(double);
    cudaFuncSetAttribute(encode_table_kernel<MPST_BASIS_TABLE_LEGENDRE_PROJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(encode_table_kernel<MPST_BASIS_TABLE_SAHAND_LEGENDRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(encode_table_kernel<MPST_BASIS_TABLE_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    switch (t.kind) {
        case MPST_BASIS_TABLE_LEGENDRE_PROJ:
            encode_table_kernel<MPST_BASIS_TABLE_LEGENDRE_PROJ><<<(unsigned)blocks, threads, smem, c->stream>>>(x, n, d, ip, dp, out, ldo);
            break;
        case MPST_BASIS_TABLE_SAHAND_LEGENDRE:
            encode_table_kernel<MPST_BASIS_TABLE_SAHAND_LEGENDRE><<<(unsigned)blocks, threads, smem, c->stream>>>(x, n, d, ip, dp, out, ldo);
            break;
        default:
            encode_table_kernel<MPST_BASIS_TABLE_SPLIT><<<(unsigned)blocks, threads, smem, c->stream>>>(x, n, d, ip, dp, out, ldo);
            break;
    }
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}
