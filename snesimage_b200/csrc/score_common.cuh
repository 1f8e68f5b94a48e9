// Shared pieces of the fused SSIMULACRA2 candidate scorers (k_score_v3, and k_score_v2 kept as its A/B check):
// the argument block, the cp.async / shared-window helpers and the pooling kernel.
//
//   error() = 100 - compute_frame_ssimulacra2(src, dst)      (lib.rs:503-548, ssimulacra2 0.5.1)
//
// One work item = one (evaluation, channel): per scale the crate blurs {i2, i2*i2, i1*i2} with a separable
// recursive Gaussian (3 second-order IIR sections, horizontal then vertical, each line a serial chain) and feeds
// mu2, s22, s12 together with the image's precomputed mu1, s11, i1 into ssim_map and edge_diff_map.  The scorers keep
// every blur plane in shared memory; what they write is six f64 sums per (evaluation, scale, channel).
#pragma once
#include "kernels.cuh"

namespace snes {

struct FusedArgs {
    const ImgDev *imgs;
    const CandEntry *cents;
    int ncand, e0, S, CS;
    int ovr;                 // >= 0: every evaluation replaces one palette entry, named by its CandEntry::slot; -1: none
    const uint8_t *maps;     // [chunk][NPIX] palette_maps of the evaluations (ignored when from_image)
    int from_image;
    int gi_fmt;              // maps hold global entry indices (k_assign_* with gi_fmt), not palette_map values
    const float *xyb_rm;     // [chunk][EVAL_XYB_FLOATS] candidate pyramid, scales >= 1 filled
    double *partials;        // [E][NSCALES][3][NSUMS]
};

// 4-byte global -> shared copy that bypasses registers (LDGSTS); sdst is a 32-bit shared-window address
__device__ __forceinline__ void cp_async4(unsigned sdst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned sdst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// k_pool_fused: Msssim::score of ssimulacra2 (108-weight pooling, cubic, power) and error() = 100 - score (lib.rs:547)
// over the scorers' partial sums [E][scale][channel][6].  One thread per evaluation.
__global__ void __launch_bounds__(128) k_pool_fused(const double *partials, int E, double *scores) {
    const int e = blockIdx.x * 128 + threadIdx.x;
    if (e >= E) return;
    const double *pe = partials + (size_t)e * (NSCALES * 3 * NSUMS);
    double ssim = 0.0;
    int i = 0;
    for (int c = 0; c < 3; c++)
        for (int s = 0; s < NSCALES; s++) {
            const int d = W >> s;
            const double opp = 1.0 / (double)(d * d);
            const double *sum = pe + ((size_t)s * 3 + c) * NSUMS;
            const double ssim0 = opp * sum[0], ssim1 = sqrt(sqrt(opp * sum[1]));
            const double e0 = opp * sum[2], e1 = sqrt(sqrt(opp * sum[3]));
            const double e2 = opp * sum[4], e3 = sqrt(sqrt(opp * sum[5]));
            ssim = fma(c_weight[i++], fabs(ssim0), ssim);
            ssim = fma(c_weight[i++], fabs(e0), ssim);
            ssim = fma(c_weight[i++], fabs(e2), ssim);
            ssim = fma(c_weight[i++], fabs(ssim1), ssim);
            ssim = fma(c_weight[i++], fabs(e1), ssim);
            ssim = fma(c_weight[i++], fabs(e3), ssim);
        }
    ssim *= 0.9562382616834844;
    ssim = fma(6.248496625763138e-5 * ssim * ssim, ssim, fma(2.326765642916932, ssim, -0.020884521182843837 * ssim * ssim));
    double score = 100.0;
    if (ssim > 0.0) score = fma(pow(ssim, 0.6276336467831387), -10.0, 100.0);
    scores[e] = 100.0 - score;
}

}  // namespace snes
