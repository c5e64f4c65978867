// FP64 issue-rate probe for sm_100a: register-only DFMA and DMMA (mma.sync m8n8k4 f64) loops.
// Calibration tool only (not on the product path). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dfma_kernel(double *out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void dmma_kernel(double *out, int iters, double a, double b) {
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { c0[i] = threadIdx.x + i; c1[i] = i; }
    double av = a + threadIdx.x * 1e-9, bv = b - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) dmma884(c0[i], c1[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int wps = 4; wps <= 32; wps *= 2) {
        int threads = wps * 32 > 1024 ? 1024 : wps * 32;
        int blocks = sms * ((wps * 32 + threads - 1) / threads);
        {
            float ms = time_it([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 8 * iters * (double)blocks * threads;
            printf("DFMA  chains=8 warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32);
            printf("DMMA  chains=8 warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 256 * 16 * iters * (double)blocks * (threads / 32);
            printf("DMMA  chains=16 warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
        }
    }
    // sustained: ~2 s of DMMA to see power-capped clocks
    {
        int threads = 512, blocks = sms;
        float ms = time_it([&] { for (int r = 0; r < 40; r++) dmma_kernel<16><<<blocks, threads>>>(out, iters * 4, 1.0000001, 1e-9); });
        double fl = 40.0 * 2.0 * 256 * 16 * iters * 4 * (double)blocks * (threads / 32);
        printf("DMMA sustained 16 warps/SM: %.1f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    {
        int threads = 512, blocks = sms;
        float ms = time_it([&] { for (int r = 0; r < 40; r++) dfma_kernel<8><<<blocks, threads>>>(out, iters * 4, 1.0000001, 1e-9); });
        double fl = 40.0 * 2.0 * 8 * iters * 4 * (double)blocks * threads;
        printf("DFMA sustained 16 warps/SM: %.1f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    return 0;
}
