// Small HBM-/latency-bound kernels around the bond GEMMs:
//   flatten_bt  (reference Training/RealRealHighDimension.jl:221-238)
//   per-sample overlap epilogue, loss + gradient weights (loss_functions.jl:302-320, 535-558)
//   TSGO / GD update and Frobenius renormalisation (loss_functions.jl:27-86, 177-179)
//   core re-layout between the wire format and the two device orientations
#include "mpst_common.cuh"

namespace {

// B[c][p + Dl*q] = sum_m Wl(s_l, a, m[, c]) * Wr(s_r, m, b[, c]);  p = s_l + d*a, q = s_r + d*b
__global__ void __launch_bounds__(256)
flatten_kernel(CoreView l, CoreView r, int d, int chi_l, int chi_m, int chi_r, int C, double* __restrict__ B) {
    const int Dl = d * chi_l, Dr = d * chi_r;
    const int64_t D = (int64_t)Dl * Dr;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * C) return;
    const int c = (int)(e / D);
    const int64_t rem = e - (int64_t)c * D;
    const int q = (int)(rem / Dl), p = (int)(rem - (int64_t)q * Dl);
    const int a = p / d, sl = p - a * d;
    const int b = q / d, sr = q - b * d;
    const double* lp = l.p + sl * l.ss + a * l.sa + c * l.sc;
    const double* rp = r.p + sr * r.ss + b * r.sb + c * r.sc;
    double s = 0.0;
    for (int m = 0; m < chi_m; m++) s += lp[m * l.sb] * rp[m * r.sa];
    B[e] = s;
}

// yhat[i] = sum_q Z[i][q] * xr[i][s_r] * R[i][b]   (one warp per sample row)
__global__ void __launch_bounds__(256)
rowdot_q_kernel(const double* __restrict__ Z, int64_t ldz, const double* __restrict__ xr,
                const double* __restrict__ R, int64_t row_begin, int64_t row_end, int d, int chi_r,
                double* __restrict__ yhat) {
    const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= row_end) return;
    const int lane = threadIdx.x & 31;
    const int Dr = d * chi_r;
    const double* z = Z + (row - row_begin) * ldz;
    const double* x = xr + row * d;
    const double* rr = R + row * chi_r;
    double s = 0.0;
    for (int q = lane; q < Dr; q += 32) {
        const int b = q / d, sr = q - b * d;
        s += z[q] * (x[sr] * rr[b]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) yhat[row] = s;
}

// yhat[i] = sum_m A[i][m] * Bm[i][m]
__global__ void __launch_bounds__(256)
rowdot_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ Bm, int64_t ldb,
              int64_t row_begin, int64_t row_end, int n, double* __restrict__ yhat) {
    const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= row_end) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int m = lane; m < n; m += 32) s += A[row * lda + m] * Bm[row * ldb + m];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) yhat[row] = s;
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    __syncthreads();
    return v;   // valid in warp 0
}

// KLD (loss_functions.jl:318, 366-371, 424-425): sample i of class c:
//   loss_i = -log(yhat^2)/denom_c,  w[c][i] = -1/(denom_c*yhat);  w[c'][i] = 0 for c' != c
// MSE (:553, :610-615): loss_i = sum_c 0.5*(yhat_c - delta)^2/N,  w[c][i] = (yhat_c - delta)/N
// partial sums per block -> red[], summed in order by final_sum_kernel.
__global__ void __launch_bounds__(256)
loss_w_kernel(const double* __restrict__ yhat, double* __restrict__ w, int64_t N, int64_t Npad, int C,
              const int64_t* __restrict__ class_off, const double* __restrict__ denom, int loss_kind,
              double* __restrict__ red, int* __restrict__ nonfinite) {
    __shared__ double sh[8];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double li = 0.0;
    if (i < N) {
        int ci = 0;
        while (ci + 1 < C && i >= class_off[ci + 1]) ci++;
        if (loss_kind == MPST_LOSS_KLD) {
            const double y = yhat[(int64_t)ci * Npad + i];
            li = -log(y * y) / denom[ci];
            for (int c = 0; c < C; c++) w[(int64_t)c * Npad + i] = (c == ci) ? -1.0 / (denom[ci] * y) : 0.0;
        } else {
            for (int c = 0; c < C; c++) {
                const double df = yhat[(int64_t)c * Npad + i] - (c == ci ? 1.0 : 0.0);
                li += 0.5 * df * df / denom[0];
                w[(int64_t)c * Npad + i] = df / denom[0];
            }
        }
    }
    // a zero / NaN overlap (KLD weight -1/(N yhat)) must not flow silently into the update and the SVD
    if (!isfinite(li)) atomicOr(nonfinite, 1);
    const double s = block_sum(li, sh);
    if (threadIdx.x == 0) red[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ red) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = v[i];
        s += t * t;
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) red[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
final_sum_kernel(const double* __restrict__ red, int n, double* __restrict__ out, int* __restrict__ nonfinite) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += red[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) {
        *out = s;
        if (!isfinite(s)) atomicOr(nonfinite, 1);
    }
}

// TSGO: B -= eta * G / ||G||   (loss_functions.jl:79);  GD: B -= eta * G   (:49)
__global__ void __launch_bounds__(256)
axpy_kernel(double* __restrict__ B, const double* __restrict__ G, int64_t n, const double* __restrict__ gnorm2,
            double eta, int tsgo) {
    const double f = tsgo ? eta / sqrt(*gnorm2) : eta;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        B[i] -= f * G[i];
}

__global__ void __launch_bounds__(256)
scale_kernel(double* __restrict__ v, int64_t n, const double* __restrict__ norm2) {
    const double f = 1.0 / sqrt(*norm2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] *= f;
}

__global__ void __launch_bounds__(256)
scale_const_kernel(double* __restrict__ v, int64_t n, double f) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] *= f;
}

// generic strided re-layout of a core: element (s,a,b,c) src strides -> dst strides
__global__ void __launch_bounds__(256)
permute_core_kernel(const double* __restrict__ src, double* __restrict__ dst, int d, int chi_l, int chi_r,
                    int C, long ss, long sa, long sb, long sc, long ds, long da, long db, long dc) {
    const int64_t n = (int64_t)d * chi_l * chi_r * C;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int64_t t = e;
    const int s = (int)(t % d); t /= d;
    const int a = (int)(t % chi_l); t /= chi_l;
    const int b = (int)(t % chi_r); t /= chi_r;
    const int c = (int)t;
    dst[s * ds + a * da + b * db + c * dc] = src[s * ss + a * sa + b * sb + c * sc];
}

// out[j][i] = in[i][j] for an (rows x cols) row-major matrix -> (cols x rows); used to move
// T x N column-major host data into site-major device layout.
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t rows,
                                 int64_t cols, int64_t ldo) {
    __shared__ double tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.y * 32, r0 = (int64_t)blockIdx.x * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t r = r0 + k, cc = c0 + threadIdx.x;
        if (r < rows && cc < cols) tile[k][threadIdx.x] = in[r * cols + cc];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t cc = c0 + k, r = r0 + threadIdx.x;
        if (r < rows && cc < cols) out[cc * ldo + r] = tile[threadIdx.x][k];
    }
}

__global__ void fill_kernel(double* __restrict__ v, int64_t n, double val) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = val;
}

// argmax_c yhat[c][i]^2 (first maximum), summary.jl:116-136; also writes yhat C x n column-major
__global__ void argmax_kernel(const double* __restrict__ yhat, int64_t Npad, int64_t n, int C,
                              double* __restrict__ out_yhat, int64_t* __restrict__ out_arg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int best = 0;
    double bv = -1.0;
    for (int c = 0; c < C; c++) {
        const double y = yhat[(int64_t)c * Npad + i];
        out_yhat[i * C + c] = y;
        if (y * y > bv) { bv = y * y; best = c; }
    }
    out_arg[i] = best;
}
// Per-sample terms of MSE_loss_acc(_conf) (summary.jl:33-114) from the overlaps y[c][i]: quadratic cost
// 0.5 sum_c (y_c - delta)^2, KL term -log(y_label^2), correct = (argmax_c |y_c| == label), confusion[label][pred]++.
// Block partial sums (fixed order -> deterministic); counts through integer atomics (exact in any order).
// labels == nullptr: sample i0 + i belongs to the class whose sorted range [class_off[c], class_off[c+1]) contains it.
__global__ void __launch_bounds__(256)
metrics_kernel(const double* __restrict__ y, int64_t ldy, int64_t n, int C, const int64_t* __restrict__ labels,
               const int64_t* __restrict__ class_off, int64_t i0, double* __restrict__ p_mse, double* __restrict__ p_kld,
               double* __restrict__ p_acc, unsigned long long* __restrict__ conf) {
    __shared__ double sh[8];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double mse = 0.0, kld = 0.0, acc = 0.0;
    if (i < n) {
        int li;
        if (labels) li = (int)labels[i];
        else { li = 0; while (li + 1 < C && i0 + i >= class_off[li + 1]) li++; }
        int pred = 0;
        double best = -1.0;
        for (int c = 0; c < C; c++) {
            const double v = y[(int64_t)c * ldy + i];
            const double df = v - (c == li ? 1.0 : 0.0);
            mse += df * df;
            if (fabs(v) > best) { best = fabs(v); pred = c; }
            if (c == li) kld = -log(v * v);
        }
        mse *= 0.5;
        acc = pred == li ? 1.0 : 0.0;
        atomicAdd(&conf[(size_t)li * C + pred], 1ull);
    }
    mse = block_sum(mse, sh);
    kld = block_sum(kld, sh);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) { p_mse[blockIdx.x] = mse; p_kld[blockIdx.x] = kld; p_acc[blockIdx.x] = acc; }
}
}  // namespace

static inline unsigned grid_for(int64_t n, int threads = 256, int64_t cap = 148 * 16) {
    int64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    return (unsigned)(g > cap ? cap : g);
}

int launch_flatten(mpst_ctx* c, CoreView l, CoreView r, int d, int chi_l, int chi_m, int chi_r, int C, double* B) {
    const int64_t n = (int64_t)d * chi_l * d * chi_r * C;
    flatten_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(l, r, d, chi_l, chi_m, chi_r, C, B);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_rowdot_q(mpst_ctx* c, const double* Z, int64_t ldz, const double* xr, const double* R,
                    int64_t row_begin, int64_t row_end, int d, int chi_r, double* yhat) {
    if (row_end <= row_begin) return MPST_OK;
    const int64_t rows = row_end - row_begin;
    rowdot_q_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, c->stream>>>(Z, ldz, xr, R, row_begin, row_end, d, chi_r, yhat);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_rowdot(mpst_ctx* c, const double* A, int64_t lda, const double* Bm, int64_t ldb, int64_t row_begin,
                  int64_t row_end, int n, double* yhat) {
    if (row_end <= row_begin) return MPST_OK;
    const int64_t rows = row_end - row_begin;
    rowdot_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, c->stream>>>(A, lda, Bm, ldb, row_begin, row_end, n, yhat);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_final_sum(mpst_ctx* c, const double* red, int n, double* out_dev) {
    final_sum_kernel<<<1, 256, 0, c->stream>>>(red, n, out_dev, c->nonfinite);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_loss_w(mpst_ctx* c, int loss_kind, const int64_t* class_off_dev, const double* denom_dev,
                  double* loss_out_dev) {
    const int blocks = (int)((c->N + 255) / 256);
    TRY(ensure_buf(c, &c->red, &c->redcap, (size_t)blocks + 4096));
    loss_w_kernel<<<blocks, 256, 0, c->stream>>>(c->yhat, c->w, c->N, c->Npad, c->C, class_off_dev, denom_dev,
                                                loss_kind, c->red, c->nonfinite);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return launch_final_sum(c, c->red, blocks, loss_out_dev);
}

int launch_sumsq(mpst_ctx* c, const double* v, int64_t n, double* out_dev) {
    const int blocks = (int)grid_for(n, 256, 1024);
    TRY(ensure_buf(c, &c->red, &c->redcap, (size_t)blocks + 4096));
    sumsq_partial_kernel<<<blocks, 256, 0, c->stream>>>(v, n, c->red);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return launch_final_sum(c, c->red, blocks, out_dev);
}

int launch_axpy(mpst_ctx* c, double* B, const double* G, int64_t n, const double* gnorm2_dev, double eta, int tsgo) {
    axpy_kernel<<<grid_for(n), 256, 0, c->stream>>>(B, G, n, gnorm2_dev, eta, tsgo);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_scale_dev(mpst_ctx* c, double* v, int64_t n, const double* norm2_dev) {
    scale_kernel<<<grid_for(n), 256, 0, c->stream>>>(v, n, norm2_dev);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_scale_const(mpst_ctx* c, double* v, int64_t n, double f) {
    scale_const_kernel<<<grid_for(n), 256, 0, c->stream>>>(v, n, f);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_permute_core(mpst_ctx* c, const double* src, double* dst, int d, int chi_l, int chi_r, int C,
                        long ss, long sa, long sb, long sc, long ds, long da, long db, long dc) {
    const int64_t n = (int64_t)d * chi_l * chi_r * C;
    permute_core_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(src, dst, d, chi_l, chi_r, C, ss, sa, sb,
                                                                          sc, ds, da, db, dc);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_transpose(mpst_ctx* c, const double* in, double* out, int64_t rows, int64_t cols, int64_t ldo) {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    dim3 block(32, 8);
    if (grid.y > 65535) { c->err = "transpose: too many columns"; return MPST_E_UNSUPPORTED; }
    transpose_kernel<<<grid, block, 0, c->stream>>>(in, out, rows, cols, ldo);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_fill(mpst_ctx* c, double* v, int64_t n, double val) {
    fill_kernel<<<grid_for(n), 256, 0, c->stream>>>(v, n, val);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_argmax(mpst_ctx* c, const double* yhat, int64_t Npad, int64_t n, int C, double* out_yhat, int64_t* out_arg) {
    argmax_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(yhat, Npad, n, C, out_yhat, out_arg);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}

int launch_metrics(mpst_ctx* c, const double* y, int64_t ldy, int64_t n, int C, const int64_t* labels_dev,
                   const int64_t* class_off_dev, int64_t i0, double* p_mse, double* p_kld, double* p_acc, int nblocks,
                   unsigned long long* conf_dev) {
    // every one of the `nblocks` partial slots is written (blocks past n contribute zeros)
    metrics_kernel<<<nblocks, 256, 0, c->stream>>>(y, ldy, n, C, labels_dev, class_off_dev, i0, p_mse, p_kld, p_acc, conf_dev);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return MPST_OK;
}
