// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue/throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_scalar(float *out, float a, float b, int iters) {
    float x[2 * ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 2 * ILP; i++) x[i] = __fmaf_rn(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_packed(float *out, float a, float b, int iters) {
    float2 x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __ffma2_rn(x[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 1024 * sizeof(float) * 4);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        const int blocks = 148 * (warps * 32 / threads);
        auto report = [&](const char *name, float ms, double flop_per_thread_iter) {
            double fl = flop_per_thread_iter * iters * (double)threads * blocks;
            printf("warps/SM %2d %-16s %8.3f ms  %7.2f TFLOP/s  (%.2f fma-lanes/clk/SM @1.965GHz)\n", warps, name, ms,
                   fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
        };
        report("scalar ILP4 (8)", timeit([&] { k_scalar<4><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); }), 16);
        report("packed ILP4 (8)", timeit([&] { k_packed<4><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); }), 16);
        report("scalar ILP2 (4)", timeit([&] { k_scalar<2><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); }), 8);
        report("packed ILP2 (4)", timeit([&] { k_packed<2><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); }), 8);
        report("packed ILP1 (2)", timeit([&] { k_packed<1><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); }), 4);
        report("scalar ILP1 (2)", timeit([&] { k_scalar<1><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); }), 4);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
