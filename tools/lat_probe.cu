// latency probe for FP64 dependent chains, rsqrt, shuffles and barriers on one SM (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, double x0) {
    __shared__ double sbuf[1024];
    double x = x0 + threadIdx.x * 1e-9, y = 0.999;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = fma(x, y, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // 2. dependent rsqrt chain
    double z = x + 1.5;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) z = rsqrt(z + 1.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // 3. shuffle + DADD chain
    double s = z;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) s += __shfl_xor_sync(0xffffffffu, s, 1 + (i & 7), 16);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // 4. barrier chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // 5. STS -> BAR -> LDS round trip
    t0 = clock64();
    double w = s;
#pragma unroll
    for (int i = 0; i < 64; i++) {
        sbuf[(threadIdx.x + i) & 1023] = w;
        __syncthreads();
        w += sbuf[(threadIdx.x + i + 17) & 1023];
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // 6. float chain
    float f = (float)w;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) f = fmaf(f, 0.999f, 1e-9f);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // 7. double sqrt, division chains
    double q = w + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) q = sqrt(q + 1.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) q = 1.0 / (q + 1.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    // 8. DFMA throughput: 8 independent chains
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) {
        a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
        a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = (t1 - t0);
    out[threadIdx.x] = x + z + s + w + f + q + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8192); cudaMalloc(&cyc, 128);
    for (int nt : {32, 256, 1024}) {
        probe<<<1, nt>>>(out, cyc, 1.0);
        probe<<<1, nt>>>(out, cyc, 1.0);
        long long h[9];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads=%4d  dfma/op %.1f  rsqrt+add/op %.1f  shfl+dadd/op %.1f  bar/op %.1f  sts-bar-lds-dadd/op %.1f  ffma/op %.1f  sqrt+add %.1f  div+add %.1f  8x-indep dfma per fma %.1f\n",
               nt, h[0] / 256.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 256.0, h[6] / 64.0, h[7] / 64.0, h[8] / 512.0);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
