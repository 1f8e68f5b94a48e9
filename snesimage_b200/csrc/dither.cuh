// k_assign_dither: optimize() with error diffusion (lib.rs:425-501), one CTA per evaluation.
//
// The reference walks the image in raster order and pushes each pixel's quantisation error, damped by
// 0.8, to E / SW / S / SE with weights 7,3,5,1 /16 in f64.  Pixel (x, y) therefore depends on
// (x-1, y) and on (x-1..x+1, y-1): it can run at wavefront step t = x + 2y, 766 steps per image with
// at most 128 rows busy.  Thread y owns row y; at step t it handles x = t - 2y, keeps the error of
// its previous pixel in registers (the E term) and a sliding window of the three errors of the row
// above, which arrive one per step through a double-buffered shared-memory mailbox.
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

template <bool LAB>
__global__ void __launch_bounds__(256) k_assign_dither(const ImgDev *imgs, const CandEntry *cents, int ncand, int e0, int S,
                                                       int CS, int ovr, uint8_t *maps, int to_image, int gi_fmt) {
    __shared__ uchar4 pal[MAX_ENTRIES];
    __shared__ float4 pal_lab[LAB ? MAX_ENTRIES : 1];
    __shared__ double mail[2][H][3];
    const int e = blockIdx.x, ea = e0 + e, img = ea / ncand, y = threadIdx.x;
    const ImgDev im = imgs[img];
    for (int j = y; j < CS; j += 256) {
        pal[j] = (j == ovr) ? cents[ea].rgb8 : im.tables->rgb8[j];
        if (LAB) {
            const float *l = (j == ovr) ? cents[ea].lab : im.tables->lab[j];
            pal_lab[j] = make_float4(l[0], l[1], l[2], 0.0f);
        }
    }
    for (int c = 0; c < 3; c++) mail[0][y][c] = mail[1][y][c] = 0.0;
    __syncthreads();

    const double w_e = 7.0 / 16.0, w_sw = 3.0 / 16.0, w_s = 5.0 / 16.0, w_se = 1.0 / 16.0, damp = 0.8;
    const uchar4 *row = im.rgba + y * W;
    uint8_t *out = (to_image ? im.map : maps + (size_t)e * NPIX) + y * W;
    const uint8_t *tp = im.tile_pal + (y >> 3) * 32;
    double ea_[3] = {0.0, 0.0, 0.0}, eb[3] = {0.0, 0.0, 0.0}, ec[3] = {0.0, 0.0, 0.0};  // row above: x-1, x, x+1
    double ee[3] = {0.0, 0.0, 0.0};                                                      // this row: x-1
    uint32_t packed = 0;

    for (int t = 0; t < W + 2 * (H - 1); t++) {
        const int x = t - 2 * y;
        if (x >= -1 && x < W) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                ea_[c] = eb[c];
                eb[c] = ec[c];
                ec[c] = (y > 0 && x + 1 < W) ? mail[(t - 1) & 1][y - 1][c] : 0.0;
            }
        }
        if (x >= 0 && x < W) {
            const uchar4 p = __ldg(row + x);
            const int sub = tp[x >> 3] * S;
            double err[3], target[3];
            const int o[3] = {p.x, p.y, p.z};
            int t8[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                double acc = 0.0;
                if (y > 0) {
                    if (x > 0) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(ea_[c], damp), w_se));
                    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(eb[c], damp), w_s));
                    if (x + 1 < W) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(ec[c], damp), w_sw));
                }
                if (x > 0) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(ee[c], damp), w_e));
                err[c] = acc;
                target[c] = __dadd_rn((double)o[c], acc);
                // lib.rs:773-778: clamp(0,255).round() as u8 (half away from zero)
                double v = target[c] < 0.0 ? 0.0 : (target[c] > 255.0 ? 255.0 : target[c]);
                t8[c] = (int)round(v);
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
                int best = 0x7fffffff;
                for (int j = 0; j < S; j++) {
                    const uchar4 cc = pal[sub + j];
                    const int key = redmean_key(cc.x, cc.y, cc.z, t8[0], t8[1], t8[2]);
                    if (key < best) {
                        best = key;
                        bi = j;
                    }
                }
            }
            const uchar4 nc = pal[sub + bi];
            if (p.w > 0) {
                ee[0] = __dsub_rn(target[0], (double)nc.x);
                ee[1] = __dsub_rn(target[1], (double)nc.y);
                ee[2] = __dsub_rn(target[2], (double)nc.z);
            } else {
                ee[0] = err[0];
                ee[1] = err[1];
                ee[2] = err[2];
                bi = 0;
            }
            mail[t & 1][y][0] = ee[0];
            mail[t & 1][y][1] = ee[1];
            mail[t & 1][y][2] = ee[2];
            packed |= (uint32_t)(gi_fmt ? (p.w > 0 ? sub + bi : GI_BLACK) : bi) << (8 * (x & 3));
            if ((x & 3) == 3) {
                *reinterpret_cast<uint32_t *>(out + (x & ~3)) = packed;
                packed = 0;
            }
        }
        __syncthreads();
    }
}

}  // namespace snes
