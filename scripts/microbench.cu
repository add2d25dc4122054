// Latency / throughput microbenchmarks that drive the forward-kernel design:
// dependent DFMA, DADD, 64-bit SHFL, LDS.64, __syncthreads.
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096

__global__ void k_dfma(double *out, double a, double b, long long *cyc)
{
    double x = threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dfma4(double *out, double a, double b, long long *cyc)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma(float *out, float a, float b, long long *cyc)
{
    float x = threadIdx.x * 1e-9f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fmaf(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(double *out, long long *cyc)
{
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_up_sync(0xffffffffu, x, 1) + 1.0;
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl32(float *out, long long *cyc)
{
    float x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_up_sync(0xffffffffu, x, 1) + 1.0f;
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds(double *out, long long *cyc)
{
    __shared__ double s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (double) ((i * 7 + 1) & 1023);
    __syncthreads();
    int idx = threadIdx.x & 1023;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) idx = (int) s[idx];
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = idx;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_bar(double *out, long long *cyc)
{
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = 0;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_div(double *out, double a, long long *cyc)
{
    double x = 1.0 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) x = a / x + 1.0;
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char *name, F f, int threads, int blocks = 1)
{
    long long *cyc;
    cudaMallocManaged(&cyc, sizeof(long long) * blocks);
    f(cyc);
    cudaDeviceSynchronize();
    f(cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-28s threads=%4d blocks=%3d : %.2f cycles/iter  (%s)\n", name, threads,
           blocks, (double) cyc[0] / N, cudaGetErrorString(e));
    cudaFree(cyc);
}

int main()
{
    double *out; float *outf;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaMalloc(&outf, sizeof(float) * 148 * 1024);
    int ths[] = { 32, 128, 320, 512, 1024 };
    for (int t : ths) {
        run("DFMA dependent chain", [&](long long *c) { k_dfma<<<1, t>>>(out, 1.0000001, 1e-9, c); }, t);
        run("DFMA 4 indep chains (per 4)", [&](long long *c) { k_dfma4<<<1, t>>>(out, 1.0000001, 1e-9, c); }, t);
    }
    run("DFMA 4 chains, full chip", [&](long long *c) { k_dfma4<<<148, 1024>>>(out, 1.0000001, 1e-9, c); }, 1024, 148);
    run("FFMA dependent chain", [&](long long *c) { k_ffma<<<1, 32>>>(outf, 1.0000001f, 1e-9f, c); }, 32);
    run("SHFL.64 + DADD dependent", [&](long long *c) { k_shfl<<<1, 32>>>(out, c); }, 32);
    run("SHFL.64 + DADD dependent", [&](long long *c) { k_shfl<<<1, 320>>>(out, c); }, 320);
    run("SHFL.32 + FADD dependent", [&](long long *c) { k_shfl32<<<1, 32>>>(outf, c); }, 32);
    run("LDS.64 pointer chase", [&](long long *c) { k_lds<<<1, 32>>>(out, c); }, 32);
    for (int t : { 64, 160, 320, 352, 1024 })
        run("__syncthreads", [&](long long *c) { k_bar<<<1, t>>>(out, c); }, t);
    run("DDIV + DADD dependent", [&](long long *c) { k_div<<<1, 32>>>(out, 3.0, c); }, 32);
    return 0;
}
