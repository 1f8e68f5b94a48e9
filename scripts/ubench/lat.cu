// Dependent-chain latencies (cycles per op, one warp) on sm_100a: FFMA, FFMA2, DFMA, DADD, DMUL, F2F, MUFU.RCP, f32 div, LDS.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void k(float *out, long long *cyc, float a, float b) {
    float x = a + threadIdx.x;
    float2 x2 = make_float2(a, b);
    double d = a, db = b, dc = 1.0 + b;
    __shared__ float sh[64];
    sh[threadIdx.x & 63] = 0.0f;
    __syncthreads();
    int idx = 0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __fmaf_rn(x, a, b);
        if (OP == 1) x2 = __ffma2_rn(x2, make_float2(a, a), make_float2(b, b));
        if (OP == 2) d = fma(d, db, dc);
        if (OP == 3) d = d + db;
        if (OP == 4) d = d * dc;
        if (OP == 5) { x = (float)d; d = (double)x + db; }          // F2F.F32.F64 + F2F.F64.F32 + DADD
        if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x));
        if (OP == 7) x = b / x;
        if (OP == 8) { idx = (int)sh[idx]; }                          // LDS + F2I
        if (OP == 9) x = x + b;
        if (OP == 10) x2 = __fadd2_rn(x2, make_float2(b, b));
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + x2.x + x2.y + (float)d + idx;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const char *names[] = {"FFMA", "FFMA2", "DFMA", "DADD", "DMUL", "F2F32+F2F64+DADD", "MUFU.RCP", "div.rn.f32", "LDS+F2I", "FADD", "FADD2"};
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1.0001f, 0.5f); k<OP><<<1, 32>>>(out, cyc, 1.0001f, 0.5f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-18s %6.1f cycles/iter\n", names[OP], (double)h / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
