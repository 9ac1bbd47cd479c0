// B200 FP64 tensor-core micro-benchmark: per-SM throughput of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) against plain DFMA, and its
// dependent-chain latency.  One DMMA = 8 x 8 x 4 = 256 FMAs per warp (8 per thread); numbers quoted in DESIGN.md §5.4.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_indep(double *out, double a, double b, int iters, long long *cycles) {
    double c[ILP][2];
#pragma unroll
    for (int k = 0; k < ILP; k++) { c[k][0] = threadIdx.x * 1e-9 + k; c[k][1] = k; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < ILP; k++) dmma(c[k][0], c[k][1], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ILP>
__global__ void dfma_indep(double *out, double a, double b, int iters, long long *cycles) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 1e-9 + k;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        dmma_indep<8><<<1, warps * 32>>>(out, 1.0000001, 1e-9, iters, cyc);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double n = (double)iters * 4 * 8 * warps;
        printf("DMMA.8x8x4  %2d warps on one SM, 8 independent accumulators per warp: %.3f warp-DMMA / clk / SM = %.1f FMA lanes / clk / SM\n", warps, n / h, n / h * 256);
    }
    dmma_indep<1><<<1, 32>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("DMMA.8x8x4  dependent chain latency: %.2f cycles\n", (double)h / (iters * 4));
    for (int warps : {4, 16, 32}) {
        dfma_indep<8><<<1, warps * 32>>>(out, 1.0000001, 1e-9, iters, cyc);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double n = (double)iters * 4 * 8 * warps;
        printf("DFMA        %2d warps on one SM, 8 independent chains per thread:      %.3f warp-DFMA / clk / SM = %.1f FMA lanes / clk / SM\n", warps, n / h, n / h * 32);
    }
    // whole GPU, to cross-check against the clock: 148 x 16 warps
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dmma_indep<8><<<148, 512>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaEventRecord(e0);
    dmma_indep<8><<<148, 512>>>(out, 1.0000001, 1e-9, iters * 4, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA whole GPU: %.2f TFLOP/s (2 flops per FMA)\n", 2.0 * 256 * (double)iters * 4 * 4 * 8 * 16 * 148 / (ms * 1e-3) / 1e12);
    dfma_indep<8><<<148, 512>>>(out, 1.0000001, 1e-9, iters * 4, cyc);
    cudaEventRecord(e0);
    dfma_indep<8><<<148, 512>>>(out, 1.0000001, 1e-9, iters * 4, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA whole GPU: %.2f TFLOP/s\n", 2.0 * 32 * (double)iters * 4 * 4 * 8 * 16 * 148 / (ms * 1e-3) / 1e12);
    return 0;
}
