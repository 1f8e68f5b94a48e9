// Candidate evaluation without dithering: optimize() (lib.rs:425-501) restricted to what a candidate can change.
//
// A candidate replaces ONE entry (slot = palette*S + index) of the image's palette
// (lib.rs:205-220, 252-262, 296-306).  Without error diffusion every pixel is decided on its own
// (lib.rs:447-451), so only pixels of tiles bound to that subpalette can change, and for those the
// reference's strict-< first-minimum over the S entries equals
//        combine( first-minimum over the entries j != index  ,  the candidate's own distance )
// with ties going to the lower index.  The first part does not depend on the candidate's colour, and it follows from
// two per-pixel records that do not depend on the replaced entry either: the pixel's best entry (first minimum over
// all S) and its runner-up (first minimum over the entries other than the best).  The first minimum over j != index
// is the best when best != index, and the runner-up otherwise.
//
//   k_assign_prepare   once per image per step: base assignment of every pixel under the current palette as a
//                      global entry index (gi), the runner-up's index and both keys.
//   k_assign_pyr       once per candidate: one distance per affected pixel -> gi map of the candidate, fused
//                      with the coarse scales (>= 1) of its XYB pyramid, which need every pixel anyway.  The replaced
//                      entry is read per evaluation (CandEntry::slot), so one launch can hold several entries' candidates.
//
// Results are identical to running the full S-entry search per candidate (tests compare both paths with
// the oracle).  RGB keys are the int32 red-mean key; Lab keys are the f32 CIEDE2000 distances (bit pattern
// stored as int).
#pragma once
#include "kernels.cuh"
#include "lab.cuh"

namespace snes {

// grid = (64, nimg), block 256, 4 horizontally adjacent pixels per thread
template <bool LAB>
__global__ void __launch_bounds__(256) k_assign_prepare(const ImgDev *imgs, int S, int CS) {
    __shared__ uchar4 pal[MAX_ENTRIES];
    __shared__ float4 pal_lab[LAB ? MAX_ENTRIES : 1];
    const ImgDev im = imgs[blockIdx.y];
    const int tid = threadIdx.x;
    for (int j = tid; j < CS; j += 256) {
        pal[j] = im.tables->rgb8[j];
        if (LAB) pal_lab[j] = make_float4(im.tables->lab[j][0], im.tables->lab[j][1], im.tables->lab[j][2], 0.0f);
    }
    __syncthreads();
    const int q = blockIdx.x * 256 + tid, px0 = q * 4, y = px0 >> 8, x = px0 & 255;
    const int sub = im.tile_pal[(y >> 3) * 32 + (x >> 3)] * S;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(im.rgba) + q);
    const uint32_t pix[4] = {v.x, v.y, v.z, v.w};
    uint32_t gi4 = 0, si4 = 0;
    int2 kk[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int r = pix[k] & 255, g = (pix[k] >> 8) & 255, b = (pix[k] >> 16) & 255, a = pix[k] >> 24;
        // one ascending pass keeps (best, runner-up) as first minima: a new best demotes the old best, which has the lowest
        // index among the entries of its key seen so far
        int bi = 0, si = 0xFF;
        if (LAB) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(im.lab) + px0 + k);
            float best = __int_as_float(0x7f800000), sec = __int_as_float(0x7f800000);
            for (int j = 0; j < S; j++) {
                const float4 c = pal_lab[sub + j];
                const float d = ciede2000(c.x, c.y, c.z, t.x, t.y, t.z);
                if (d < best) {
                    sec = best;
                    si = j == 0 ? 0xFF : bi;
                    best = d;
                    bi = j;
                } else if (d < sec) {
                    sec = d;
                    si = j;
                }
            }
            kk[k] = make_int2(__float_as_int(best), __float_as_int(sec));
        } else {
            int best = 0x7fffffff, sec = 0x7fffffff;
            for (int j = 0; j < S; j++) {
                const uchar4 c = pal[sub + j];
                const int key = redmean_key(c.x, c.y, c.z, r, g, b);
                if (key < best) {
                    sec = best;
                    si = j == 0 ? 0xFF : bi;
                    best = key;
                    bi = j;
                } else if (key < sec) {
                    sec = key;
                    si = j;
                }
            }
            kk[k] = make_int2(best, sec);
        }
        gi4 |= (uint32_t)(a > 0 ? sub + bi : GI_BLACK) << (8 * k);
        si4 |= (uint32_t)si << (8 * k);
    }
    reinterpret_cast<uint32_t *>(im.base_gi)[q] = gi4;
    reinterpret_cast<uint32_t *>(im.sec_idx)[q] = si4;
    reinterpret_cast<int4 *>(im.keys)[2 * q] = make_int4(kk[0].x, kk[0].y, kk[1].x, kk[1].y);
    reinterpret_cast<int4 *>(im.keys)[2 * q + 1] = make_int4(kk[2].x, kk[2].y, kk[3].x, kk[3].y);
}

// k_assign_pyr: grid = (4, evaluations of the chunk), block 256; a CTA owns one 128x128 quadrant.
//   MODE 0 / 1: delta assignment with the red-mean key / CIEDE2000 (writes the candidate's gi map)
//   MODE 2    : the gi map already exists (written by k_assign_dither); only the pyramid is built
// and in every mode the row-major XYB planes of scales 1..5 (the fused scorer's inputs): a thread takes
// 4x4 scale-0 pixels -> 2x2 pixels of scale 1 and one of scale 2; scales 3..5 go through shared memory
// (downscale_by_2 on linear RGB, then linear_rgb_to_xyb + make_positive_xyb, as ssimulacra2 does).
//
// base_xyb (MODE 0 / 1, optional): the coarse pyramids of the images' prepared base assignment, [image][EVAL_XYB_FLOATS].
// A 4x4 block in which the candidate changes no assignment and no pixel uses the replaced entry has the base image's
// scale-1 and scale-2 pixels, so they are
// copied instead of recomputed (three f64-Halley cube roots per pixel); the pixels of the blocks that did change are
// queued in shared memory and converted by all threads of the CTA afterwards, so that the few changed blocks of a
// candidate do not serialise whole warps.
constexpr int PYR_MAXJOBS = 1020;   // a multiple of the 5 pixels a block queues, so a refused block leaves no gap

// MODE 0 (one integer key per affected pixel) is bound by the latency of its dependent loads: 64 registers for a fourth CTA
// per SM pay (0.67 -> 0.59 ms per 4096 evaluations); the CIEDE2000 and full-pyramid modes are faster with their registers
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 0 ? 4 : 3) k_assign_pyr(const ImgDev *imgs, const CandEntry *cents, int ncand, int e0, int S, int CS,
                                                    int has_ovr, uint8_t *maps, float *xyb_rm_base, const float *base_xyb) {
    __shared__ float s_lin[MAX_ENTRIES + 1][3];
    __shared__ float4 s_jobs[(MODE == 2) ? 1 : PYR_MAXJOBS];   // linear RGB + output offset of a queued pixel
    __shared__ int s_njobs;
    __shared__ uint16_t s_queue[MODE == 1 ? 8 : 1][MODE == 1 ? 512 : 1];   // MODE 1: per warp, the pixels whose distance must be computed
    __shared__ uint32_t s_take[MODE == 1 ? 8 : 1][32];                       // MODE 1: per warp and lane, the 16 decisions of the lane's block
    __shared__ float s2[3][32][33];   // linear RGB of the quadrant at scale 2
    __shared__ float s3[3][16][17];
    __shared__ float s4[3][8][9];
    const int e = blockIdx.y, ea = e0 + e, img = ea / ncand, tid = threadIdx.x;
    const int qx0 = (blockIdx.x & 1) * 128, qy0 = (blockIdx.x >> 1) * 128;
    const ImgDev im = imgs[img];
    float *rm = xyb_rm_base + (size_t)e * EVAL_XYB_FLOATS;
    uint8_t *map = maps + (size_t)e * NPIX;
    const CandEntry ce = cents[ea];
    const int ovr = has_ovr >= 0 ? ce.slot : -1;   // the entry this evaluation replaces
    for (int j = tid; j < CS; j += 256)
        for (int c = 0; c < 3; c++) s_lin[j][c] = (j == ovr) ? ce.lin[c] : im.tables->lin[j][c];
    if (tid < 3) {
        s_lin[BLACK][tid] = im.tables->lin[BLACK][tid];
        s_lin[GI_BLACK][tid] = im.tables->lin[BLACK][tid];   // C*S <= 255 on this path: slot 255 is free
    }
    const int psub = ovr >= 0 ? (ovr / S) * S : 0, oloc = ovr - psub;
    const int cr = ce.rgb8.x, cg = ce.rgb8.y, cb = ce.rgb8.z;
    // bytes of a gi word that lie in the replaced entry's subpalette [psub, psub + S): the pixels this candidate can change
    // (GI_BLACK = 255 is never in range: C*S <= 255 on this path)
    const uint32_t lo4 = (uint32_t)psub * 0x01010101u, hi4 = (uint32_t)(psub + S) * 0x01010101u;
    auto in_sub = [&](uint32_t g) { return __vcmpgeu4(g, lo4) & __vcmpltu4(g, hi4); };
    const float *base = (MODE != 2 && base_xyb) ? base_xyb + (size_t)img * EVAL_XYB_FLOATS : nullptr;
    const uint32_t ovr_rep = (uint32_t)ovr * 0x01010101u;
    if (tid == 0) s_njobs = 0;
    __syncthreads();

    auto store_xyb = [&](int L, int x, int y, const float lin[3]) {
        const int d = W >> L;
        float xv, yv, bv;
        lin_to_pxyb(lin[0], lin[1], lin[2], xv, yv, bv);
        const size_t o = 3 * (size_t)scale_off(L) + (size_t)y * d + x;
        rm[o] = xv;
        rm[o + (size_t)d * d] = yv;
        rm[o + 2 * (size_t)d * d] = bv;
    };

    // ---- scales 0 (gi map), 1 and 2: 32x32 blocks of 4x4 pixels, 4 blocks per thread -----------------------------------------
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
        const int bx = tid & 31, by = it * 8 + (tid >> 5);   // block inside the quadrant
        const int x0 = qx0 + 4 * bx, y0 = qy0 + 4 * by;
        uint32_t g4[4], a4[4];   // gi words of the block's rows; a4: 0xFF in the bytes of pixels the candidate can change
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int px = (y0 + r) * W + x0;
            if (MODE == 2) {
                g4[r] = __ldg(reinterpret_cast<const uint32_t *>(map + px));
                a4[r] = 0u;
            } else {
                g4[r] = __ldg(reinterpret_cast<const uint32_t *>(im.base_gi + px));
                a4[r] = in_sub(g4[r]);
            }
        }
        // (best, runner-up) of pixel px against the candidate's own distance: the first minimum over the other entries is the
        // best unless the best IS the replaced entry, then the runner-up
        auto others = [&](int px, int bi, int &xi, int &xkey) {
            const int2 kk = __ldg(im.keys + px);
            xi = bi;
            xkey = kk.x;
            if (bi == oloc) {
                xi = im.sec_idx[px];
                xkey = kk.y;
            }
        };
        // MODE 1: CIEDE2000 is expensive and only ~1/C of the blocks are affected, so the warp pools them.  Two passes:
        //   A  two affected blocks at a time, one pixel per lane: a lower bound of the candidate's distance from the lightness
        //      term alone decides most pixels without the formula (below); the pixels it cannot decide are queued per warp;
        //   B  the queue, 32 pixels at a time: the full distance; a "take" sets the pixel's bit in its owner's word.
        // Lower bound: dE00^2 = a^2 + b^2 + c^2 + RT b c with a = dL'/SL, b = dC'/SC, c = dH'/SH and |RT| <= 2 sqrt(x / (x + 25^7)) < 2,
        // so b^2 + c^2 + RT b c >= (1 - |RT|/2)(b^2 + c^2) >= 0 and dE00 >= |dL|/SL; SL = 1 + 0.015 m / sqrt(20 + m), m = (Lbar - 50)^2
        // <= 2500, is at most 1.747.  A candidate whose |dL| / 1.75 exceeds the pixel's threshold (the key of the best of the other
        // entries) by more than the margin below cannot win or tie, whatever the f32 evaluation of the formula rounds to.
        uint32_t takebits = 0;
        if (MODE == 1) {
            const int lane = tid & 31, wp = tid >> 5;
            uint16_t *queue = s_queue[wp];
            s_take[wp][lane] = 0u;
            __syncwarp();
            const bool aff = (a4[0] | a4[1] | a4[2] | a4[3]) != 0u;
            unsigned mask = __ballot_sync(0xffffffffu, aff);
            int nq = 0;
            while (mask) {
                const int sa = __ffs(mask) - 1;
                mask &= mask - 1;
                const bool hasb = mask != 0;
                const int sb = hasb ? __ffs(mask) - 1 : sa;
                if (hasb) mask &= mask - 1;
                const int src = lane < 16 ? sa : sb;
                const int pix = lane & 15, r = pix >> 2, c = pix & 3;
                const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
                const uint32_t w0 = __shfl_sync(0xffffffffu, g4[0], src), w1 = __shfl_sync(0xffffffffu, g4[1], src);
                const uint32_t w2 = __shfl_sync(0xffffffffu, g4[2], src), w3 = __shfl_sync(0xffffffffu, g4[3], src);
                const uint32_t w = r == 0 ? w0 : (r == 1 ? w1 : (r == 2 ? w2 : w3));
                const int gi = (w >> (8 * c)) & 255;
                bool undecided = false;
                const int px = (sy0 + r) * W + sx0 + c;
                if (gi >= psub && gi < psub + S && (lane < 16 || hasb)) {
                    int xi, xkey;
                    others(px, gi - psub, xi, xkey);
                    const float xb = __int_as_float(xkey);
                    const float tl = __ldg(im.lab + 4 * (size_t)px);   // L of the pixel
                    const float lb = fabsf(ce.lab[0] - tl) * (1.0f / 1.75f);
                    undecided = !(lb > __fmaf_rn(xb, 1.001f, 1e-3f));   // (NaN or infinite thresholds stay undecided)
                }
                const unsigned ub = __ballot_sync(0xffffffffu, undecided);
                if (undecided) queue[nq + __popc(ub & ((1u << lane) - 1u))] = (uint16_t)px;
                nq += __popc(ub);
            }
            __syncwarp();
            for (int sidx = lane; sidx < nq; sidx += 32) {
                const int px = queue[sidx], py = px >> 8, pxx = px & 255;
                const int gi = im.base_gi[px];
                const float4 t = __ldg(reinterpret_cast<const float4 *>(im.lab) + px);
                const float d = ciede2000(ce.lab[0], ce.lab[1], ce.lab[2], t.x, t.y, t.z);
                int xi, xkey;
                others(px, gi - psub, xi, xkey);
                const float xb = __int_as_float(xkey);
                if (d < xb || (d == xb && oloc < xi))
                    atomicOr(&s_take[wp][(pxx >> 2) & 31], 1u << (4 * (py & 3) + (pxx & 3)));   // owner lane = the block's column
            }
            __syncwarp();
            takebits = s_take[wp][lane];
            __syncwarp();
        }
        float l1[2][2][3];
        uint32_t outw[4];
        bool changed = false;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            uint32_t out4 = g4[r];
            if (MODE != 2 && a4[r] != 0u) {
                out4 = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    int gi = (g4[r] >> (8 * c)) & 255;
                    if ((a4[r] >> (8 * c)) & 1u) {   // affected tile, opaque pixel: candidate entry vs the best of the others
                        const int px = (y0 + r) * W + x0 + c;
                        int xi, xkey;
                        others(px, gi - psub, xi, xkey);
                        bool take;
                        if (MODE == 1) {
                            take = (takebits >> (4 * r + c)) & 1;
                        } else {
                            const uchar4 p = __ldg(im.rgba + px);
                            const int key = redmean_key(cr, cg, cb, p.x, p.y, p.z);
                            take = key < xkey || (key == xkey && oloc < xi);
                        }
                        gi = psub + (take ? oloc : xi);
                    }
                    out4 |= (uint32_t)gi << (8 * c);
                }
            }
            if (MODE != 2) *reinterpret_cast<uint32_t *>(map + (y0 + r) * W + x0) = out4;
            {   // the block differs from the base image if an assignment changed or a pixel sits on the replaced entry
                const uint32_t v = out4 ^ ovr_rep;
                changed |= out4 != g4[r] || ((v - 0x01010101u) & ~v & 0x80808080u) != 0;
            }
            outw[r] = out4;
        }
        float l2[3];
        const bool copy = base && !changed;
        if (copy) {
            // an unchanged block has the base image's linear scale-2 pixel as well (k_pyramid keeps it in the unused scale-0 area
            // of the base pyramid buffer): no table lookups, no downscaling
#pragma unroll
            for (int c = 0; c < 3; c++) l2[c] = __ldg(base + c * 4096 + (y0 >> 2) * 64 + (x0 >> 2));
        } else {
            float quad[4][4][3];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int gi = (outw[r] >> (8 * c)) & 255;
                    quad[r][c][0] = s_lin[gi][0];
                    quad[r][c][1] = s_lin[gi][1];
                    quad[r][c][2] = s_lin[gi][2];
                }
            // downscale_by_2: ((p00 + p01) + p10) + p11, * 0.25  (row-major 2x2 order of the crate's loop)
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        l1[a][b][c] = (((quad[2 * a][2 * b][c] + quad[2 * a][2 * b + 1][c]) + quad[2 * a + 1][2 * b][c]) +
                                       quad[2 * a + 1][2 * b + 1][c]) * 0.25f;
#pragma unroll
            for (int c = 0; c < 3; c++) l2[c] = (((l1[0][0][c] + l1[0][1][c]) + l1[1][0][c]) + l1[1][1][c]) * 0.25f;
        }
        const size_t o1 = 3 * (size_t)scale_off(1) + (size_t)(y0 >> 1) * 128 + (x0 >> 1);   // scale 1: 128 x 128
        const size_t o2 = 3 * (size_t)scale_off(2) + (size_t)(y0 >> 2) * 64 + (x0 >> 2);    // scale 2: 64 x 64
        int slot = -1;
        if (base && changed) {
            slot = atomicAdd(&s_njobs, 5);
            if (slot + 5 > PYR_MAXJOBS) slot = -1;   // queue full: convert in place
        }
        if (copy) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const size_t oc = o1 + (size_t)c * 128 * 128;
                *reinterpret_cast<float2 *>(rm + oc) = __ldg(reinterpret_cast<const float2 *>(base + oc));
                *reinterpret_cast<float2 *>(rm + oc + 128) = __ldg(reinterpret_cast<const float2 *>(base + oc + 128));
                rm[o2 + (size_t)c * 64 * 64] = __ldg(base + o2 + (size_t)c * 64 * 64);
            }
        } else if (slot >= 0) {
            s_jobs[slot + 0] = make_float4(l1[0][0][0], l1[0][0][1], l1[0][0][2], __int_as_float((int)o1));
            s_jobs[slot + 1] = make_float4(l1[0][1][0], l1[0][1][1], l1[0][1][2], __int_as_float((int)o1 + 1));
            s_jobs[slot + 2] = make_float4(l1[1][0][0], l1[1][0][1], l1[1][0][2], __int_as_float((int)o1 + 128));
            s_jobs[slot + 3] = make_float4(l1[1][1][0], l1[1][1][1], l1[1][1][2], __int_as_float((int)o1 + 129));
            s_jobs[slot + 4] = make_float4(l2[0], l2[1], l2[2], __int_as_float(-(int)o2 - 1));   // negative: a scale-2 pixel
        } else {
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) store_xyb(1, (x0 >> 1) + b, (y0 >> 1) + a, l1[a][b]);
            store_xyb(2, x0 >> 2, y0 >> 2, l2);
        }
        s2[0][by][bx] = l2[0];
        s2[1][by][bx] = l2[1];
        s2[2][by][bx] = l2[2];
    }
    __syncthreads();
    // ---- the queued pixels of the blocks the candidate changed
    if (MODE != 2) {
        const int nj = s_njobs < PYR_MAXJOBS ? s_njobs : PYR_MAXJOBS;
        for (int j = tid; j < nj; j += 256) {
            const float4 jb = s_jobs[j];
            const int oi = __float_as_int(jb.w);
            const size_t o = oi >= 0 ? (size_t)oi : (size_t)(-oi - 1);
            const size_t plane = oi >= 0 ? (size_t)128 * 128 : (size_t)64 * 64;
            float xv, yv, bv;
            lin_to_pxyb(jb.x, jb.y, jb.z, xv, yv, bv);
            rm[o] = xv;
            rm[o + plane] = yv;
            rm[o + 2 * plane] = bv;
        }
    }
    // ---- scale 3: 16x16 per quadrant, one pixel per thread ------------------------------------------------------------------
    {
        const int ax = tid & 15, ay = tid >> 4;
        float l3[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            l3[c] = (((s2[c][2 * ay][2 * ax] + s2[c][2 * ay][2 * ax + 1]) + s2[c][2 * ay + 1][2 * ax]) + s2[c][2 * ay + 1][2 * ax + 1]) * 0.25f;
        store_xyb(3, (qx0 >> 3) + ax, (qy0 >> 3) + ay, l3);
        s3[0][ay][ax] = l3[0];
        s3[1][ay][ax] = l3[1];
        s3[2][ay][ax] = l3[2];
    }
    __syncthreads();
    if (tid < 64) {   // scale 4: 8x8
        const int ax = tid & 7, ay = tid >> 3;
        float l4[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            l4[c] = (((s3[c][2 * ay][2 * ax] + s3[c][2 * ay][2 * ax + 1]) + s3[c][2 * ay + 1][2 * ax]) + s3[c][2 * ay + 1][2 * ax + 1]) * 0.25f;
        store_xyb(4, (qx0 >> 4) + ax, (qy0 >> 4) + ay, l4);
        s4[0][ay][ax] = l4[0];
        s4[1][ay][ax] = l4[1];
        s4[2][ay][ax] = l4[2];
    }
    __syncthreads();
    if (tid < 16) {   // scale 5: 4x4
        const int ax = tid & 3, ay = tid >> 2;
        float l5[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            l5[c] = (((s4[c][2 * ay][2 * ax] + s4[c][2 * ay][2 * ax + 1]) + s4[c][2 * ay + 1][2 * ax]) + s4[c][2 * ay + 1][2 * ax + 1]) * 0.25f;
        store_xyb(5, (qx0 >> 5) + ax, (qy0 >> 5) + ay, l5);
    }
}

}  // namespace snes
