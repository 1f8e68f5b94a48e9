// k_assign_dither: optimize() with error diffusion (lib.rs:425-501), one CTA per evaluation.
//
// The reference walks the image in raster order and pushes each pixel's quantisation error, damped by
// 0.8, to E / SW / S / SE with weights 7,3,5,1 /16 in f64.  Pixel (x, y) therefore depends on
// (x-1, y) and on (x-1..x+1, y-1): it can run at wavefront step t = x + 2y, 766 steps per image with
// at most 128 rows busy.  Thread i owns rows i and i + 128 (the second row starts exactly when the first one ends, so a
// thread is busy for 512 of the 766 steps); at step t it handles pixel tau = t - 2i of its rows, keeps the error of its
// previous pixel in registers (the E term) and a sliding window of the three errors of the row above, which arrive one
// per step from the thread above through a shared-memory mailbox and one block barrier per step.  (A hand-over by warp
// shuffles and a ring between warps was measured slower: profiles/r2_dither_ab.txt.)
//
// The step itself is dc::step (dither_core.h, also compiled for the host by tests/test_dither_core.py).  What this
// generation changes against round 2's first kernel (4.61 -> 3.64 ms per 4096 evaluations, profiles/r2_dither_ab.txt) is the instruction count
// of a step -- the kernel runs at IPC 2.2-2.4 with eight CTAs per SM, so instructions are what it pays for:
//   * the loop is unrolled by three and the window's roles rotate with the step number: no register moves;
//     the mailbox is a ring of three buffers addressed by the same step number: no pointer swaps;
//   * packed keys (dither_core.h): v = 8 key' + entry number as one int32, minimum by VIMNMX -- 8 instead of 10 instructions
//     per palette entry (2 LDS, 5 IMAD, 1 VIMNMX);
//   * round-half-away of a target as two round-down additions instead of F2I / I2F / DADD / DSETP / select, entry colours
//     kept as doubles: no conversion instruction is left on the chain from the mailbox read to the mailbox write.
//
// Bit-exactness: the accumulated error of a pixel is built in the reference's += order, i.e. the
// raster order of the contributing pixels -- SE term from (x-1,y-1), S from (x,y-1), SW from
// (x+1,y-1), then E from (x-1,y) -- each term being (value * 0.8) * weight, with no FMA contraction
// (__dmul_rn/__dadd_rn).  Transparent pixels forward their accumulated error unchanged and get index 0
// (lib.rs:453-475).  Distances: the red-mean integer key (common.cuh) or CIEDE2000 (lab.cuh).
#pragma once
#include "common.cuh"
#include "dither_core.h"
#include "lab.cuh"

namespace snes {

constexpr int DITHER_THREADS = dc::THREADS;
#ifndef DITHER_MIN_CTAS
#define DITHER_MIN_CTAS 8   // 64 registers per thread
#endif
static_assert(dc::IW == W && dc::IH == H && dc::GI_TRANSPARENT == GI_BLACK, "dither_core.h restates common.cuh");

// dynamic shared memory of k_assign_dither<LAB> for a palette of CS entries
inline size_t dither_smem_bytes(int CS, bool lab) {
    return sizeof(double) * 9 * DITHER_THREADS + NTILES + (size_t)CS * (sizeof(dc::PalD) + (lab ? sizeof(float4) : sizeof(dc::KeyCoef)));
}

// grid = evaluations, block = DITHER_THREADS.
template <bool LAB>
__global__ void __launch_bounds__(DITHER_THREADS, DITHER_MIN_CTAS) k_assign_dither(const ImgDev *imgs, const CandEntry *cents, int ncand, int e0, int S,
                                                                  int CS, int has_ovr, uint8_t *maps, int to_image, int gi_fmt,
                                                                  const TileMove *moves /* per evaluation, or null */) {
    // dynamic shared memory, sized by the palette (dither_smem_bytes): at 8 x 15 colours 17.7 KB instead of 26.6, so that the
    // register file, not shared memory, decides how many CTAs an SM holds
    extern __shared__ __align__(32) unsigned char dsm[];
    double(*mail)[3][DITHER_THREADS] = reinterpret_cast<double(*)[3][DITHER_THREADS]>(dsm);   // [step % 3][channel][thread]: conflict-free 8-byte accesses
    uint8_t *s_tp = dsm + sizeof(double) * 9 * DITHER_THREADS;                                 // first entry of each tile's subpalette, in dc::stp_slot order
    dc::PalD *pald = reinterpret_cast<dc::PalD *>(s_tp + NTILES);                              // as_rgba(entry) as doubles
    dc::KeyCoef *ktab = reinterpret_cast<dc::KeyCoef *>(pald + CS);                            // RGB: packed red-mean key coefficients per entry
    float4 *pal_lab = reinterpret_cast<float4 *>(pald + CS);                                   // LAB: Lab of the entries (same place)
    const int e = blockIdx.x, ea = e0 + e, img = ea / ncand, i = threadIdx.x;
    const ImgDev im = imgs[img];
    const int ovr = has_ovr >= 0 ? cents[ea].slot : -1;   // the entry this evaluation replaces (per evaluation: CandEntry::slot)
    for (int j = i; j < CS; j += DITHER_THREADS) {
        const uchar4 c = (j == ovr) ? cents[ea].rgb8 : im.tables->rgb8[j];
        dc::PalD p;
        p.v[0] = (double)c.x;
        p.v[1] = (double)c.y;
        p.v[2] = (double)c.z;
        p.v[3] = 0.0;
        pald[j] = p;
        if (!LAB) ktab[j] = dc::key_coef(c.x, c.y, c.z, j % S);
        if (LAB) {
            const float *l = (j == ovr) ? cents[ea].lab : im.tables->lab[j];
            pal_lab[j] = make_float4(l[0], l[1], l[2], 0.0f);
        }
    }
    for (int j = i; j < NTILES; j += DITHER_THREADS)
        s_tp[dc::stp_slot(j)] = (uint8_t)(((moves && moves[ea].tile == j) ? moves[ea].sub : im.tile_pal[j]) * S);
    for (int c = 0; c < 9; c++) (&mail[0][0][0])[c * DITHER_THREADS + i] = 0.0;
    __syncthreads();

    uint8_t *out_row = (to_image ? im.map : maps + (size_t)e * NPIX) + i * W;
    const uint4 *src_row = reinterpret_cast<const uint4 *>(im.rgba + i * W);
    const int up = (i + DITHER_THREADS - 1) & (DITHER_THREADS - 1);  // the thread that owns the row above
    auto load_quad = [&](int off, uint32_t (&q)[4]) {
        const uint4 v = __ldg(src_row + (off >> 2));
        q[0] = v.x;
        q[1] = v.y;
        q[2] = v.z;
        q[3] = v.w;
    };
    auto nearest = [&](int first, int r, int g, int b) -> int {
        if (LAB) {
            float tl, ta, tb;
            srgb8_to_lab(r, g, b, tl, ta, tb);
            float best = __int_as_float(0x7f800000);
            int bi = 0;
            for (int j = 0; j < S; j++) {
                const float4 cl = pal_lab[first + j];
                const float d = ciede2000(cl.x, cl.y, cl.z, tl, ta, tb);
                if (d < best) {
                    best = d;
                    bi = j;
                }
            }
            return bi;
        } else {
            return dc::nearest_rgb(ktab + first, S, r, g, b);
        }
    };
    dc::Thread th;
    dc::thread_init(th);
    load_quad(0, th.q);
    int tau = -2 * i;
    const uint32_t gi_mask = gi_fmt ? 0xffu : 0u;
#pragma unroll 1
    for (int t = 0; t < dc::STEPS; t += 3, tau += 3) {
        dc::step<0>(th, tau, i, &mail[2][0][up], &mail[0][0][i], s_tp, pald, gi_mask, out_row, nearest, load_quad);
        __syncthreads();
        dc::step<1>(th, tau + 1, i, &mail[0][0][up], &mail[1][0][i], s_tp, pald, gi_mask, out_row, nearest, load_quad);
        __syncthreads();
        dc::step<2>(th, tau + 2, i, &mail[1][0][up], &mail[2][0][i], s_tp, pald, gi_mask, out_row, nearest, load_quad);
        __syncthreads();
    }
}

}  // namespace snes
