// Micro-benchmark: FP64 / FP32 FMA issue rate, dependent-issue latency and 64-bit shuffle latency on this GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peak pipe_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void k_fma(T* out, T a, T b, int iters) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// all three operands are per-thread registers (no uniform / constant operand)
template <int ILP>
__global__ void k_fma3(double* out, const double* in, int iters) {
    double acc[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 64 + i]; c[i] = in[threadIdx.x + 128 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], b[i], c[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// two register operands + one shared operand (b[i] * v + acc[i]) like a biquad update with a common input
template <int ILP>
__global__ void k_fma2(double* out, const double* in, int iters) {
    double acc[ILP], b[ILP];
    double v = in[threadIdx.x];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 64 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(b[i], v, acc[i]);
        v = acc[0];
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_add2(double* out, const double* in, int iters) {
    double acc[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 64 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] + b[i];
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K>
double time_kernel(K launch, double ops) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ops / (ms * 1e-3);
}
__global__ void k_lat_dfma(double* out, double a, double b, int iters, long long* cyc) {
    double acc = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc = fma(acc, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lat_shfl(double* out, int iters, long long* cyc) {
    double acc = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc = __shfl_up_sync(0xffffffffu, acc, 1);
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <typename T, int ILP>
double run(int blocks, int threads, int iters) {
    T* out; cudaMalloc(&out, sizeof(T) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_fma<T, ILP><<<blocks, threads>>>(out, (T)1.0000001, (T)1e-9, iters);
    cudaEventRecord(e0);
    k_fma<T, ILP><<<blocks, threads>>>(out, (T)1.0000001, (T)1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(out);
    return (double)blocks * threads * iters * ILP / (ms * 1e-3);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int warps = 1; warps <= 16; warps *= 2) {
        double r64 = run<double, 8>(p.multiProcessorCount * 4, warps * 32 / 4 < 32 ? 32 : warps * 32 / 4, 20000);
        printf("fp64 ILP8 threads/block=%d x4 blocks/SM: %.2f T FMA/s\n", warps * 32 / 4 < 32 ? 32 : warps * 32 / 4, r64 / 1e12);
    }
    printf("fp64 peak (1024 thr/SM, ILP8): %.2f T FMA/s\n", run<double, 8>(p.multiProcessorCount * 2, 512, 20000) / 1e12);
    printf("fp64 peak (512 thr/SM, ILP4): %.2f T FMA/s\n", run<double, 4>(p.multiProcessorCount * 2, 256, 40000) / 1e12);
    printf("fp64 (128 thr/SM, ILP1): %.2f T FMA/s\n", run<double, 1>(p.multiProcessorCount, 128, 100000) / 1e12);
    printf("fp64 (128 thr/SM, ILP2): %.2f T FMA/s\n", run<double, 2>(p.multiProcessorCount, 128, 100000) / 1e12);
    printf("fp64 (128 thr/SM, ILP4): %.2f T FMA/s\n", run<double, 4>(p.multiProcessorCount, 128, 100000) / 1e12);
    printf("fp32 peak (1024 thr/SM, ILP8): %.2f T FMA/s\n", run<float, 8>(p.multiProcessorCount * 2, 512, 40000) / 1e12);
    {
        double *o, *in; cudaMalloc(&o, 8 * 148 * 2 * 512); cudaMalloc(&in, 8 * 2048); cudaMemset(in, 0, 8 * 2048);
        const int it = 20000; const double n = 148.0 * 2 * 512 * it * 8;
        printf("fp64 DFMA 3 register operands: %.2f T/s\n", time_kernel([&] { k_fma3<8><<<148 * 2, 512>>>(o, in, it); }, n) / 1e12);
        printf("fp64 DFMA 2 register operands + 1 shared register: %.2f T/s\n", time_kernel([&] { k_fma2<8><<<148 * 2, 512>>>(o, in, it); }, n) / 1e12);
        printf("fp64 DADD 2 register operands: %.2f T/s\n", time_kernel([&] { k_add2<8><<<148 * 2, 512>>>(o, in, it); }, n) / 1e12);
    }
    double* out; long long* cyc; cudaMalloc(&out, 8 * 32); cudaMallocManaged(&cyc, 8);
    k_lat_dfma<<<1, 32>>>(out, 1.0000001, 1e-9, 1000, cyc); cudaDeviceSynchronize();
    printf("dependent DFMA latency: %.2f cycles\n", (double)*cyc / 64000.0);
    k_lat_shfl<<<1, 32>>>(out, 1000, cyc); cudaDeviceSynchronize();
    printf("dependent 64-bit shfl_up latency: %.2f cycles\n", (double)*cyc / 64000.0);
    return 0;
}
