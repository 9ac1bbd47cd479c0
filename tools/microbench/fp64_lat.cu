// B200 FP64 micro-benchmark: dependent-chain latency and per-SM throughput of DFMA, MUFU.RCP64H and LDS.64
// (numbers quoted in DESIGN.md; nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dep_chain(double *out, double a, double b, int iters, long long *cycles) {
    double x = threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ILP>
__global__ void indep(double *out, double a, double b, int iters, long long *cycles) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 1e-9 + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
__global__ void rcp_chain(double *out, int iters, long long *cycles) {
    double x = 1.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            double r;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));      // SASS: MUFU.RCP64H
            x = r + 1.0;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
__global__ void lds_chain(double *out, int iters, long long *cycles) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 33) & 1023;
    __syncthreads();
    int j = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) j = idx[j];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = j;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    dep_chain<<<1, 32>>>(out, 1.0000001, 1e-9, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %.2f cycles\n", (double)h / (iters * 16.0));
    indep<8><<<1, 32>>>(out, 1.0000001, 1e-9, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 1 warp, ILP 8: %.2f cycles per warp-instruction\n", (double)h / (iters * 32.0));
    for (int warps : {1, 2, 4, 8, 16}) {
        indep<8><<<1, 32 * warps>>>(out, 1.0000001, 1e-9, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA %2d warps on one SM, ILP 8: %.2f warp-instructions per cycle per SM\n", warps, iters * 32.0 * warps / (double)h);
    }
    rcp_chain<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RCP64H + DADD dependent pair: %.2f cycles\n", (double)h / (iters * 8.0));
    lds_chain<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS dependent latency: %.2f cycles\n", (double)h / (iters * 16.0));
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
