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

// bfxb[px] = (|i1 - mu1| of X, |i1 - mu1| of B) at scale 0: the same f32 expression the maps evaluate per pixel.  grid 256 x 256.
__global__ void __launch_bounds__(256) k_make_bfxb(ImgDev im) {
    const int px = blockIdx.x * 256 + threadIdx.x;
    im.bfxb[px] = make_float2(fabsf(im.xyb_rm[px] - im.mu1[px]), fabsf(im.xyb_rm[2 * NPIX + px] - im.mu1[2 * NPIX + px]));
}

// Msssim::score of ssimulacra2 (108-weight pooling, cubic, power) and error() = 100 - score (lib.rs:547) for one
// evaluation's partial sums pe[scale][channel][6], by one warp: lane q < 18 prepares the six terms of (channel q / 6, scale
// q % 6) -- the loads and the fourth roots, which is where a single thread spent its time -- and every lane then runs the
// crate's accumulation over the 108 terms in the crate's order, fetching each term by shuffle, so the result is bit for bit
// what the sequential loop gives.  All 32 lanes must call it; all return the same value.
__device__ __forceinline__ double pool_score_warp(const double *pe, int lane) {
    double t[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (lane < 3 * NSCALES) {
        const int c = lane / NSCALES, s = lane - c * NSCALES, d = W >> s;
        const double opp = 1.0 / (double)(d * d);
        const double *sum = pe + ((size_t)s * 3 + c) * NSUMS;
        t[0] = fabs(opp * sum[0]);               // ssim: mean of d
        t[1] = fabs(opp * sum[2]);               // edge: mean artifact
        t[2] = fabs(opp * sum[4]);               // edge: mean detail_lost
        t[3] = fabs(sqrt(sqrt(opp * sum[1])));   // 4-norms of the same three
        t[4] = fabs(sqrt(sqrt(opp * sum[3])));
        t[5] = fabs(sqrt(sqrt(opp * sum[5])));
    }
    double ssim = 0.0;
    int i = 0;
    for (int q = 0; q < 3 * NSCALES; q++)
#pragma unroll
        for (int k = 0; k < 6; k++) ssim = fma(c_weight[i++], __shfl_sync(0xffffffffu, t[k], q), ssim);
    ssim *= 0.9562382616834844;
    ssim = fma(6.248496625763138e-5 * ssim * ssim, ssim, fma(2.326765642916932, ssim, -0.020884521182843837 * ssim * ssim));
    double score = 100.0;
    if (ssim > 0.0) score = fma(pow(ssim, 0.6276336467831387), -10.0, 100.0);
    return 100.0 - score;
}

// k_pool_fused: error() of E evaluations from their partial sums; one warp per evaluation.  grid = ceil(E / 4), block 128.
__global__ void __launch_bounds__(128) k_pool_fused(const double *partials, int E, double *scores) {
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= E) return;
    const double v = pool_score_warp(partials + (size_t)e * PART_DOUBLES, lane);
    if (lane == 0) scores[e] = v;
}

// k_pool_argmin: the end of a candidate step in one launch instead of four: error() of image j itself from its own partial
// sums (-> its current error and self_scores[j]; skipped when self_partials is null), error() of its ncand candidate
// evaluations, and their strict-< first minimum (lib.rs:216) as (error, idx_base + k).  grid = nimg, block 1024: 32 warps,
// one evaluation per warp at a time.
__global__ void __launch_bounds__(1024) k_pool_argmin(const ImgDev *imgs, const double *self_partials, double *self_scores,
                                                     const double *partials, int ncand, int idx_base, double *scores, Best *best) {
    __shared__ double s_e[32];
    __shared__ int s_i[32];
    const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double be = __longlong_as_double(0x7ff0000000000000ll);
    int bi = 0x7fffffff;
    for (int k = warp; k < ncand; k += 32) {   // ascending k per warp: a strict < keeps the first minimum
        const double v = pool_score_warp(partials + ((size_t)j * ncand + k) * PART_DOUBLES, lane);
        if (lane == 0) scores[(size_t)j * ncand + k] = v;
        if (v < be) {
            be = v;
            bi = k;
        }
    }
    if (self_partials && warp == 31) {   // the warp with the fewest candidates
        const double v = pool_score_warp(self_partials + (size_t)j * PART_DOUBLES, lane);
        if (lane == 0) {
            *imgs[j].cur_err = v;
            self_scores[j] = v;
        }
    }
    if (lane == 0) {
        s_e[warp] = be;
        s_i[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
        be = s_e[lane];
        bi = s_i[lane];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const double oe = __shfl_xor_sync(0xffffffffu, be, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (best_less(oe, oi, be, bi)) {
                be = oe;
                bi = oi;
            }
        }
        if (lane == 0) {
            Best b;
            b.err = be;
            b.idx = bi == 0x7fffffff ? -1 : bi + idx_base;
            b.pad = 0;
            best[j] = b;
        }
    }
}

// k_finish_iterate: what follows the pooling of the candidates' scores (k_pool_fused) in the speculative single-call
// iterations (snes_image_iterate), one CTA per image instead of four more launches: error() of the image itself (-> its
// current error), the strict-< first minimum of every step's list (lib.rs:216), then the steps in order up to the first
// one whose best beats the current error (k_apply_first_accept's rule; a NES step always takes its first minimum, and ends
// the run only if that changes the entry's colour).  grid = nimg, block 128.
__global__ void __launch_bounds__(128) k_finish_iterate(const ImgDev *imgs, const double *self_partials /* null: NES, no error() first */,
                                                       const double *scores, const int *step_slot, int nsteps,
                                                       const uint8_t *cand, int ncand, int force, Best *best, int *consumed, int *chosen,
                                                       double *err_before, double *err_after) {
    __shared__ double s_e[4];
    __shared__ int s_i[4];
    const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, per_img = nsteps * ncand;
    const ImgDev im = imgs[j];
    if (self_partials && warp == 0) {
        const double v = pool_score_warp(self_partials + (size_t)j * PART_DOUBLES, lane);
        if (lane == 0) {
            *im.cur_err = v;
            err_before[j] = v;
        }
    }
    const double *s_score = scores + (size_t)j * per_img;
    __syncthreads();
    int used = nsteps, pick = -1;
    for (int s = 0; s < nsteps; s++) {
        double be = __longlong_as_double(0x7ff0000000000000ll);
        int bi = 0x7fffffff;
        for (int k = tid; k < ncand; k += 128) {
            const double v = s_score[s * ncand + k];
            if (best_less(v, k, be, bi)) {
                be = v;
                bi = k;
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const double oe = __shfl_xor_sync(0xffffffffu, be, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (best_less(oe, oi, be, bi)) {
                be = oe;
                bi = oi;
            }
        }
        __syncthreads();   // s_e / s_i of the previous step have been read
        if (lane == 0) {
            s_e[warp] = be;
            s_i[warp] = bi;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 4; w++)
                if (best_less(s_e[w], s_i[w], be, bi)) {
                    be = s_e[w];
                    bi = s_i[w];
                }
            Best b;
            b.err = be;
            b.idx = bi == 0x7fffffff ? -1 : bi;
            b.pad = 0;
            best[(size_t)j * nsteps + s] = b;
            if (pick < 0 && used == nsteps && b.idx >= 0) {
                const uint8_t *c = cand + (((size_t)j * nsteps + s) * ncand + b.idx) * 3;
                const int slot = step_slot[s];
                const uint8_t *cur = im.palette + 3 * slot;
                if (force && c[0] == cur[0] && c[1] == cur[1] && c[2] == cur[2]) {
                    // NES iteration (lib.rs:242-284) whose first minimum is the colour the entry already has: the state stays what
                    // the later steps were evaluated against, and its error() is this evaluation's score
                    *im.cur_err = b.err;
                    err_before[j] = b.err;
                } else if (c[0] <= 32 && c[1] <= 32 && c[2] <= 32 && (force || b.err < *im.cur_err)) {
                    im.palette[3 * slot] = c[0];
                    im.palette[3 * slot + 1] = c[1];
                    im.palette[3 * slot + 2] = c[2];
                    *im.cur_err = b.err;
                    used = s + 1;
                    pick = s * ncand + b.idx;
                }
            }
        }
    }
    if (tid == 0) {
        consumed[j] = used;
        chosen[j] = pick;
        err_after[j] = *im.cur_err;
    }
}

}  // namespace snes
