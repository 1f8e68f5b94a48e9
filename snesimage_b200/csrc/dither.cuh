// k_assign_dither: optimize() with error diffusion (lib.rs:425-501), one CTA per evaluation.
//
// The reference walks the image in raster order and pushes each pixel's quantisation error, damped by
// 0.8, to E / SW / S / SE with weights 7,3,5,1 /16 in f64.  Pixel (x, y) therefore depends on
// (x-1, y) and on (x-1..x+1, y-1): it can run at wavefront step t = x + 2y, 766 steps per image with
// at most 128 rows busy.  Thread y owns row y; at step t it handles x = t - 2y, keeps the error of
// its previous pixel in registers (the E term) and a sliding window of the three errors of the row
// above, which arrive one per step from the thread above: by warp shuffle inside a warp, through a small
// shared-memory ring with publish / consume counters between warps (DITHER_SHFL, the default), so that no block
// barrier is left in the 766-step loop and the four warps may drift apart by up to DITHER_RING steps.  The previous
// hand-over -- a double-buffered mailbox for all 128 threads and one __syncthreads() per step -- is kept under
// DITHER_SHFL=0 for A/B runs.
//
// Bit-exactness: the accumulated error of a pixel is built in the reference's += order, i.e. the
// raster order of the contributing pixels -- SE term from (x-1,y-1), S from (x,y-1), SW from
// (x+1,y-1), then E from (x-1,y) -- each term being (value * 0.8) * weight, with no FMA contraction
// (__dmul_rn/__dadd_rn).  Transparent pixels forward their accumulated error unchanged and get index 0
// (lib.rs:453-475).  Distances: the red-mean integer key (common.cuh) or CIEDE2000 (lab.cuh).
#pragma once
#include "common.cuh"
#include "lab.cuh"

namespace snes {

constexpr int DITHER_THREADS = 128;
#ifndef DITHER_SHFL
#define DITHER_SHFL 0   // A/B on B200 (profiles/r2_dither_ab.txt): shuffles + ring 6.12 ms per 4096 evaluations at 8x15, mailbox + barrier 4.61
#endif
constexpr int DITHER_RING = 16;      // steps a producing warp may run ahead of the warp that consumes its last row
#ifndef DITHER_MIN_CTAS
#define DITHER_MIN_CTAS 6   // 80 registers per thread (ms per 4096 evaluations at 8x15: 8 CTAs = 64 registers 5.75, 7 = 72: 5.05, 6 = 80: 4.62)
#endif

// grid = evaluations, block = DITHER_THREADS.  Thread i owns row i and then row i + 128: in wavefront time
// tau = t - 2i it handles pixel x = tau of row i for tau in [0, 255] and pixel x = tau - 256 of row i + 128 for tau in
// [256, 511] -- the second row starts exactly when the first one ends, so a thread is busy for 512 of the 766 steps
// (a thread-per-row block would be busy for 256).  The row above the thread's current row is always the current row
// of thread i - 1 (thread 127 for the first pixel of row 128), so the mailbox is indexed by thread.
template <bool LAB>
__global__ void __launch_bounds__(DITHER_THREADS, DITHER_MIN_CTAS) k_assign_dither(const ImgDev *imgs, const CandEntry *cents, int ncand, int e0, int S,
                                                                  int CS, int has_ovr, uint8_t *maps, int to_image, int gi_fmt,
                                                                  const TileMove *moves /* per evaluation, or null */) {
    __shared__ int4 pal[MAX_ENTRIES];  // r, g, b of as_rgba(entry), 1024 + r
    // red-mean key of entry (R, G, B) against target (r, g, b) with the terms that depend on the target alone dropped
    // (they shift every key of a pixel by the same amount, so the strict-< first-minimum is unchanged; int32 arithmetic
    // is exact modulo 2^32 and both key and key'' fit):
    //   key   = (1024 + r + R)(R - r)^2 + 2048 (G - g)^2 + (1534 - r - R)(B - b)^2              (common.cuh)
    //   key'' = key - (r^3 + 1024 r^2 + 2048 g^2 + 1534 b^2 - r b^2)
    //         = C0 + A r - R (r^2 + b^2) - 4096 G g + C1 b + 2B (r b)
    //   C0 = (1024 + R) R^2 + 2048 G^2 + (1534 - R) B^2,  A = -R^2 - 2048 R - B^2,  C1 = -2B (1534 - R)
    // five multiply-adds per entry instead of three differences, three squares and three weighted sums.
    __shared__ int4 kc0[LAB ? 1 : MAX_ENTRIES];  // C0, A, -R, -4096 G
    __shared__ int2 kc1[LAB ? 1 : MAX_ENTRIES];  // C1, 2B
    __shared__ float4 pal_lab[LAB ? MAX_ENTRIES : 1];
    __shared__ uint8_t s_tp[NTILES];   // tile_palettes * S
#if DITHER_SHFL
    // hand-over between warps: thread 32b+31 publishes its damped error of every step into ring[b], thread (32b+32) % 128
    // consumes it one step later.  published[b] / consumed[b] count the values written / read (value of step s in slot s % RING).
    __shared__ double ring[DITHER_THREADS / 32][DITHER_RING][3];
    __shared__ int published[DITHER_THREADS / 32], consumed[DITHER_THREADS / 32];
#else
    __shared__ double mail[2][3][DITHER_THREADS];  // [buffer][channel][thread]: conflict-free 8-byte accesses
#endif
    const int e = blockIdx.x, ea = e0 + e, img = ea / ncand, i = threadIdx.x;
    const ImgDev im = imgs[img];
    const int ovr = has_ovr >= 0 ? cents[ea].slot : -1;   // the entry this evaluation replaces (per evaluation: CandEntry::slot)
    for (int j = i; j < CS; j += DITHER_THREADS) {
        const uchar4 c = (j == ovr) ? cents[ea].rgb8 : im.tables->rgb8[j];
        pal[j] = make_int4(c.x, c.y, c.z, 1024 + c.x);
        if (!LAB) {
            const int R = c.x, G = c.y, B = c.z;
            kc0[j] = make_int4((1024 + R) * R * R + 2048 * G * G + (1534 - R) * B * B, -R * R - 2048 * R - B * B, -R, -4096 * G);
            kc1[j] = make_int2(-2 * B * (1534 - R), 2 * B);
        }
        if (LAB) {
            const float *l = (j == ovr) ? cents[ea].lab : im.tables->lab[j];
            pal_lab[j] = make_float4(l[0], l[1], l[2], 0.0f);
        }
    }
    for (int j = i; j < NTILES; j += DITHER_THREADS)
        s_tp[j] = (uint8_t)(((moves && moves[ea].tile == j) ? moves[ea].sub : im.tile_pal[j]) * S);
#if DITHER_SHFL
    if (i < DITHER_THREADS / 32) published[i] = consumed[i] = 0;
#else
    for (int c = 0; c < 3; c++) mail[0][c][i] = mail[1][c][i] = 0.0;
#endif
    __syncthreads();

    const double w_e = 7.0 / 16.0, w_sw = 3.0 / 16.0, w_s = 5.0 / 16.0, w_se = 1.0 / 16.0, damp = 0.8;
    uint8_t *outb = to_image ? im.map : maps + (size_t)e * NPIX;
    const int up = (i + DITHER_THREADS - 1) & (DITHER_THREADS - 1);  // the thread that owns the row above
    double ea_[3] = {0.0, 0.0, 0.0}, eb[3] = {0.0, 0.0, 0.0}, ec[3] = {0.0, 0.0, 0.0};  // row above: x-1, x, x+1
    double ee[3] = {0.0, 0.0, 0.0};                                                      // this row: x-1 (all damped)
    uint32_t packed = 0;
    // four source pixels at a time, requested four steps before their first use
    uint4 nextq = __ldg(reinterpret_cast<const uint4 *>(im.rgba + i * W));
    uint4 curq = nextq;
#if DITHER_SHFL
    const int lane = i & 31, wid = i >> 5, b_in = (wid + DITHER_THREADS / 32 - 1) & (DITHER_THREADS / 32 - 1);
    volatile int *v_pub = published, *v_con = consumed;
    (void)up;
#else
    const double *mrd = &mail[1][0][up];  // read side of step t: buffer (t - 1) & 1
    double *mwr = &mail[0][0][i];         // write side of step t: buffer t & 1
#endif

    for (int t = 0; t < W + 2 * (H - 1); t++) {
        const int tau = t - 2 * i;
#if DITHER_SHFL
        // the thread above computed pixel x+1 of its row in the previous step: its damped error is still in its `ee`
        double upv[3];
#pragma unroll
        for (int c = 0; c < 3; c++) upv[c] = __shfl_up_sync(0xffffffffu, ee[c], 1);
        if (lane == 0 && t > 0) {   // ... unless it sits in another warp: take the value it published for step t - 1
            for (int spin = 0; v_pub[b_in] < t; spin++)
                if (spin > (1 << 24)) __trap();   // a lost hand-over must not hang the GPU
            __threadfence_block();
            const double *slot = ring[b_in][(t - 1) & (DITHER_RING - 1)];
            upv[0] = slot[0];
            upv[1] = slot[1];
            upv[2] = slot[2];
            __threadfence_block();
            v_con[b_in] = t;
        }
#endif
        if (tau >= -1 && tau < 2 * W) {
            // the row above published its pixel x+1 in the previous step (for x = 255 this is already pixel 0 of the row
            // above the thread's second row; the x+1 term of pixel 255 is weighted out below).  Row 0 has no row above:
            // its window stays zero, so the three terms from above need no y > 0 test.
            const bool has_up = i > 0 || tau >= W - 1;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                ea_[c] = eb[c];
                eb[c] = ec[c];
#if DITHER_SHFL
                ec[c] = has_up ? upv[c] : 0.0;
#else
                ec[c] = has_up ? mrd[c * DITHER_THREADS] : 0.0;
#endif
            }
        }
        if (tau >= 0 && tau < 2 * W) {
            const int x = tau & (W - 1), second = tau >> 8;  // second = 1 on the thread's row i + 128
            if ((x & 3) == 0) {
                curq = nextq;
                const int tn = tau + 4;
                if (tn < 2 * W) nextq = __ldg(reinterpret_cast<const uint4 *>(im.rgba + (i + (tn >> 8) * DITHER_THREADS) * W + (tn & (W - 1))));
            }
            const uint32_t pw = (x & 3) == 0 ? curq.x : (x & 3) == 1 ? curq.y : (x & 3) == 2 ? curq.z : curq.w;
            const int sub = s_tp[((i >> 3) + second * (DITHER_THREADS / 8)) * 32 + (x >> 3)];
            double err[3], target[3];
            const int o[3] = {(int)(pw & 255u), (int)((pw >> 8) & 255u), (int)((pw >> 16) & 255u)};
            const bool opaque = (pw >> 24) != 0;
            int t8[3];
            // A term the reference skips (lib.rs:478-493: x + 1 < width, x > 0) gets weight zero here: it contributes
            // +-0, which leaves the running sum unchanged, and the sum's leading `0.0 +` only ever changes the sign of a
            // zero, which nothing downstream can see (targets are sums with an integer, errors are only ever added).
            const double wse = x > 0 ? w_se : 0.0, wsw = x + 1 < W ? w_sw : 0.0, we = x > 0 ? w_e : 0.0;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                // every stored error is already damped (e * 0.8, the first product of each term of lib.rs:479-493); the
                // terms are added in the raster order of the contributing pixels: SE, S, SW, then E
                double acc = __dmul_rn(ea_[c], wse);
                acc = __dadd_rn(acc, __dmul_rn(eb[c], w_s));
                acc = __dadd_rn(acc, __dmul_rn(ec[c], wsw));
                acc = __dadd_rn(acc, __dmul_rn(ee[c], we));
                err[c] = acc;
                target[c] = __dadd_rn((double)o[c], acc);
                // lib.rs:773-778: clamp(0,255).round() as u8 (half away from zero).  round() is monotone and fixes 0 and
                // 255, so clamping after rounding gives the same byte; a negative target ends at 0 whatever its fraction,
                // so only the non-negative case needs the exact rule: truncate, then step up when the (exact) fraction
                // reaches one half.
                const int tz = __double2int_rz(target[c]);
                const int tr = tz + (__dsub_rn(target[c], (double)tz) >= 0.5 ? 1 : 0);
                t8[c] = min(max(tr, 0), 255);
            }
            int bi = 0;
            if (LAB) {
                float tl, ta, tb;
                srgb8_to_lab(t8[0], t8[1], t8[2], tl, ta, tb);
                float best = __int_as_float(0x7f800000);
                for (int j = 0; j < S; j++) {
                    const float4 cl = pal_lab[sub + j];
                    const float d = ciede2000(cl.x, cl.y, cl.z, tl, ta, tb);
                    if (d < best) {
                        best = d;
                        bi = j;
                    }
                }
            } else {
                // red-mean key without its target-only terms (see kc0 / kc1 above)
                const int r = t8[0], g = t8[1], b = t8[2], s2 = r * r + b * b, rb = r * b;
                int best = 0x7fffffff;
#define DITHER_TRY(j)                                                                                   \
    {                                                                                                   \
        const int4 k0 = kc0[sub + (j)];                                                                 \
        const int2 k1 = kc1[sub + (j)];                                                                 \
        const int key = k0.x + k0.y * r + k0.z * s2 + k0.w * g + k1.x * b + k1.y * rb;                  \
        if (key < best) {                                                                               \
            best = key;                                                                                 \
            bi = (j);                                                                                   \
        }                                                                                               \
    }
                if (S == 15) {  // the SNES subpalette: straight-line, entry numbers as immediates
#pragma unroll
                    for (int j = 0; j < 15; j++) DITHER_TRY(j)
                } else {
#pragma unroll 4
                    for (int j = 0; j < S; j++) DITHER_TRY(j)
                }
#undef DITHER_TRY
            }
            const int4 nc = pal[sub + bi];
            if (opaque) {
                ee[0] = __dmul_rn(__dsub_rn(target[0], (double)nc.x), damp);
                ee[1] = __dmul_rn(__dsub_rn(target[1], (double)nc.y), damp);
                ee[2] = __dmul_rn(__dsub_rn(target[2], (double)nc.z), damp);
            } else {
                ee[0] = __dmul_rn(err[0], damp);
                ee[1] = __dmul_rn(err[1], damp);
                ee[2] = __dmul_rn(err[2], damp);
                bi = 0;
            }
#if !DITHER_SHFL
            mwr[0] = ee[0];
            mwr[DITHER_THREADS] = ee[1];
            mwr[2 * DITHER_THREADS] = ee[2];
#endif
            packed |= (uint32_t)(gi_fmt ? (opaque ? sub + bi : GI_BLACK) : bi) << (8 * (x & 3));
            if ((x & 3) == 3) {
                *reinterpret_cast<uint32_t *>(outb + (i + second * DITHER_THREADS) * W + (x & ~3)) = packed;
                packed = 0;
            }
        }
#if DITHER_SHFL
        if (lane == 31) {   // publish this step's value for the first thread of the next warp
            for (int spin = 0; t - v_con[wid] >= DITHER_RING; spin++)   // slot t % RING still holds an unread value
                if (spin > (1 << 24)) __trap();
            double *slot = ring[wid][t & (DITHER_RING - 1)];
            slot[0] = ee[0];
            slot[1] = ee[1];
            slot[2] = ee[2];
            __threadfence_block();
            v_pub[wid] = t + 1;
        }
#else
        // swap the mailbox buffers for the next step
        const double *nr = mwr - i + up;
        mwr = const_cast<double *>(mrd) - up + i;
        mrd = nr;
        __syncthreads();
#endif
    }
}

}  // namespace snes
