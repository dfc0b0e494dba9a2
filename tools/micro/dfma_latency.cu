// Microbenchmark: issue cadence of FP64 FMA / FP32 FMA / MUFU rsqrt with ONE warp per SM sub-partition and
// 1..8 independent dependency chains per thread (the regime of the persistent thread-per-problem solver).
// nvcc -gencode arch=compute_100a,code=sm_100a -o build/dfma_latency tools/micro/dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, class T>
__global__ void chain(T* out, long long* cyc, int iters, T a, T b)
{
    T x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = (T)(threadIdx.x + c);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    T s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void chain_rsqrt(double* out, long long* cyc, int iters)
{
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = 1.5 + threadIdx.x + c;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = rsqrt(x[c]) + 1.5;
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH, class T>
void run(const char* name, int threads)
{
    T* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(T)); cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 2000;
    chain<CH, T><<<148, threads>>>(out, cyc, iters, (T)0.999999, (T)1e-9);
    chain<CH, T><<<148, threads>>>(out, cyc, iters, (T)0.999999, (T)1e-9);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double per = (double)h[0] / (iters * 16.0 * CH);
    printf("%s threads/SM %4d chains %d: %.2f cycles per warp-instruction (%.2f per dependent step)\n", name, threads, CH, per, per * CH);
    cudaFree(out); cudaFree(cyc);
}
template <int CH>
void run_rsqrt(int threads)
{
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 2000;
    chain_rsqrt<CH><<<148, threads>>>(out, cyc, iters);
    chain_rsqrt<CH><<<148, threads>>>(out, cyc, iters);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double per = (double)h[0] / (iters * 4.0 * CH);
    printf("rsqrt(double)+add threads/SM %4d chains %d: %.1f cycles each (%.1f per dependent step)\n", threads, CH, per, per * CH);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<1, double>("DFMA", 128); run<2, double>("DFMA", 128); run<4, double>("DFMA", 128); run<8, double>("DFMA", 128);
    run<1, double>("DFMA", 256); run<4, double>("DFMA", 256); run<8, double>("DFMA", 256);
    run<1, float>("FFMA", 128); run<2, float>("FFMA", 128); run<4, float>("FFMA", 128); run<8, float>("FFMA", 128);
    run_rsqrt<1>(128); run_rsqrt<2>(128); run_rsqrt<4>(128);
    return 0;
}
