// Probe which tensor-map box shapes the TMA unit accepts (sm_100a): one box copy into shared memory, checked on the host.
// usage: tma_probe <elem: 1|4> <dim0> <dim1> <box0> <box1> <x> <y> <dst_offset_bytes>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                           const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap *gtm, int use_global, int x, int y, unsigned bytes, unsigned off, unsigned char *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(sm + off);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        const CUtensorMap *p = use_global ? gtm : &tm;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d), "l"(p), "r"(x),
                     "r"(y), "r"(b)
                     : "memory");
    }
    unsigned ok = 0;
    for (int spin = 0; !ok && spin < (1 << 20); spin++)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b) : "memory");
    if (threadIdx.x == 0 && !ok) printf("timeout\n");
    for (unsigned i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[off + i];
}
int main(int argc, char **argv) {
    const int es = atoi(argv[1]);
    const uint64_t d0 = atoll(argv[2]), d1 = atoll(argv[3]);
    const uint32_t b0 = atoi(argv[4]), b1 = atoi(argv[5]);
    const int x = atoi(argv[6]), y = atoi(argv[7]);
    const unsigned off = atoi(argv[8]);
    const int use_global = argc > 9 ? atoi(argv[9]) : 0;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    enc_fn enc = (enc_fn)p;
    std::vector<unsigned char> h(d0 * d1 * es);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)(i * 7 + 3);
    unsigned char *g, *out;
    cudaMalloc(&g, h.size());
    cudaMemcpy(g, h.data(), h.size(), cudaMemcpyHostToDevice);
    const unsigned bytes = b0 * b1 * es;
    cudaMalloc(&out, bytes);
    CUtensorMap tm;
    cuuint64_t gd[2] = {d0, d1}, gs[1] = {d0 * es};
    cuuint32_t bx[2] = {b0, b1}, est[2] = {1, 1};
    CUresult r = enc(&tm, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, g, gd, gs, bx, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d; ", (int)r);
    CUtensorMap *gtm;
    cudaMalloc(&gtm, sizeof(tm));
    cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<1, 128, 100 * 1024>>>(tm, gtm, use_global, x, y, bytes, off, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s; ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<unsigned char> o(bytes);
        cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (uint32_t r1 = 0; r1 < b1; r1++)
            for (uint32_t c = 0; c < b0 * es; c++) {
                const long gy = (long)y + r1, gx = (long)x * es + c;
                const unsigned char want = (gy < 0 || gy >= (long)d1 || gx < 0 || gx >= (long)(d0 * es)) ? 0 : h[gy * d0 * es + gx];
                bad += o[r1 * b0 * es + c] != want;
            }
        printf("mismatches %zu of %u", bad, bytes);
    }
    printf("\n");
    return 0;
}
