// Device kernels of libsnesgpu (sm_100a) around the scorer: palette tables, the full-search assignment, the XYB
// pyramid, the source-side blur planes built once per image, argmin / merge / accept.
//   candidate step: k_tables -> k_assign_* (assign_delta.cuh, dither.cuh) -> k_score_v3 -> k_pool_fused -> k_argmin
// All of it is HBM/L2-bound byte, integer and f32 stencil work; no stage is a dense contraction,
// so no tensor cores.  Each kernel cites the reference lines (lib.rs) or the crate routine it covers.
#pragma once
#include "common.cuh"
#include "lab.cuh"

namespace snes {

__constant__ float c_lin_lut[256];   // yuvxyb sRGB EOTF of v/255, v = 0..255 (host-built with libm powf)
__constant__ float c_n2[3];          // recursive-Gaussian feed-forward taps  (libjxl CreateRecursiveGaussian)
__constant__ float c_d1[3];          // recursive-Gaussian feedback taps
__constant__ double c_weight[108];   // SSIMULACRA2 pooling weights
__constant__ uint8_t c_nes[NES_COUNT][4];  // lib.rs:687-742

// ------------------------------------------------------------------------------------------------
// k_tables: per image, expand every palette entry (SnesColor::as_rgba, lib.rs:662-669) to rgb8,
// linear RGB and positive XYB; per evaluation, do the same for the one candidate colour.
// grid = nimg + ceil(E/256) blocks of 256 threads.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fill_entry(uint8_t r5, uint8_t g5, uint8_t b5, uchar4 &rgb8, float lin[3], float xyb[3]) {
    rgb8 = snes_as_rgba(r5, g5, b5);
    lin[0] = c_lin_lut[rgb8.x];
    lin[1] = c_lin_lut[rgb8.y];
    lin[2] = c_lin_lut[rgb8.z];
    lin_to_pxyb(lin[0], lin[1], lin[2], xyb[0], xyb[1], xyb[2]);
}

// A candidate colour component above 32 (possible only in a device-resident list; host lists are refused at the
// boundary) is clamped to 32 for the tables and raises *fault, which the next snes_ctx_synchronize() reports.
// Evaluation e = (image e / ncand, candidate e % ncand) takes colour cand[(image * cand_stride + cand_lo + candidate)]:
// a rank evaluating the slice [cand_lo, cand_lo + ncand) of a list of cand_stride candidates per image reads it in place.
// The entry evaluation (j, k) replaces is slots[k] when a per-candidate slot list is given (one launch holding the
// candidates of several entries), else `slot` for all of them.
__global__ void __launch_bounds__(256) k_tables(const ImgDev *imgs, int nimg, int CS, const uint8_t *cand, int E, int ncand,
                                                int cand_stride, int cand_lo, int slot, const int *slots /* [ncand] per-candidate, or null */,
                                                CandEntry *cents,
                                                const float4 *labtab /* null unless perceptual */, int *fault) {
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < nimg) {
        const ImgDev im = imgs[blockIdx.x];
        PalTables *t = im.tables;
        if (tid < CS) {
            uchar4 rgb8;
            float lin[3], xyb[3];
            fill_entry(im.palette[3 * tid], im.palette[3 * tid + 1], im.palette[3 * tid + 2], rgb8, lin, xyb);
            t->rgb8[tid] = rgb8;
            for (int c = 0; c < 3; c++) {
                t->lin[tid][c] = lin[c];
                t->xyb[tid][c] = xyb[c];
            }
            if (labtab) {
                const float4 l = labtab[bgr555_index(im.palette[3 * tid], im.palette[3 * tid + 1], im.palette[3 * tid + 2])];
                t->lab[tid][0] = l.x;
                t->lab[tid][1] = l.y;
                t->lab[tid][2] = l.z;
                t->lab[tid][3] = 0.0f;
            }
        }
        if (tid == 0) {  // transparent pixels render as (0,0,0) (lib.rs:551-558, 570-572)
            float x, y, b;
            const float z = c_lin_lut[0];
            lin_to_pxyb(z, z, z, x, y, b);
            t->lin[BLACK][0] = t->lin[BLACK][1] = t->lin[BLACK][2] = z;
            t->xyb[BLACK][0] = x;
            t->xyb[BLACK][1] = y;
            t->xyb[BLACK][2] = b;
        }
    } else {
        const int e = ((int)blockIdx.x - nimg) * 256 + tid;
        if (e < E) {
            CandEntry ce;
            const uint8_t *cp = cand + 3 * ((size_t)(e / ncand) * cand_stride + cand_lo + (e % ncand));
            uint8_t c5[3] = {cp[0], cp[1], cp[2]};
            if (c5[0] > 32 || c5[1] > 32 || c5[2] > 32) {
                atomicOr(fault, 1);
                for (int c = 0; c < 3; c++) c5[c] = c5[c] > 32 ? 32 : c5[c];
            }
            fill_entry(c5[0], c5[1], c5[2], ce.rgb8, ce.lin, ce.xyb);
            ce.lab[0] = ce.lab[1] = ce.lab[2] = 0.0f;
            if (labtab) {
                const float4 l = labtab[bgr555_index(c5[0], c5[1], c5[2])];
                ce.lab[0] = l.x;
                ce.lab[1] = l.y;
                ce.lab[2] = l.z;
            }
            ce.slot = slots ? slots[e % ncand] : slot;
            ce.pad = 0;
            cents[e] = ce;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_assign_rgb: optimize() (lib.rs:425-501) without dithering and with the red-mean metric:
// target == original pixel, nearest entry of the tile's subpalette, strict-< first minimum
// (lib.rs:780-792), transparent pixels get index 0 (lib.rs:453-458).
// grid = (64, E), block 256; each thread owns 4 horizontally adjacent pixels (one 16-byte load,
// one 4-byte store).  to_image != 0 writes into the image's own palette_map (E == nimg).
// ------------------------------------------------------------------------------------------------
// gi_fmt != 0 (scratch maps of the fused scorer): each byte is the global entry index tile_sub*S + index, or
// GI_BLACK for a transparent pixel, so the scorer needs neither tile_palettes nor alpha to render the pixel.
__global__ void __launch_bounds__(256) k_assign_rgb(const ImgDev *imgs, const CandEntry *cents, int ncand, int e0, int S,
                                                    int CS, int has_ovr, uint8_t *maps, int to_image, int gi_fmt,
                                                    const TileMove *moves /* per evaluation, or null */) {
    __shared__ uchar4 pal[MAX_ENTRIES];
    const int e = blockIdx.y, ea = e0 + e, img = ea / ncand, tid = threadIdx.x;
    const ImgDev im = imgs[img];
    const int ovr = has_ovr >= 0 ? cents[ea].slot : -1;   // the entry this evaluation replaces (per evaluation: CandEntry::slot)
    for (int j = tid; j < CS; j += 256) pal[j] = (j == ovr) ? cents[ea].rgb8 : im.tables->rgb8[j];
    __syncthreads();
    const int q = blockIdx.x * 256 + tid;  // quad-of-4 index
    const int px0 = q * 4, y = px0 >> 8, x = px0 & 255;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(im.rgba) + q);
    const int tile = (y >> 3) * 32 + (x >> 3);
    const int sub = ((moves && moves[ea].tile == tile) ? moves[ea].sub : im.tile_pal[tile]) * S;
    const uint32_t pix[4] = {v.x, v.y, v.z, v.w};
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int r = pix[k] & 255, g = (pix[k] >> 8) & 255, b = (pix[k] >> 16) & 255, a = pix[k] >> 24;
        int best = 0x7fffffff, bi = 0;
        for (int j = 0; j < S; j++) {
            const uchar4 c = pal[sub + j];
            const int key = redmean_key(c.x, c.y, c.z, r, g, b);
            if (key < best) {
                best = key;
                bi = j;
            }
        }
        packed |= (uint32_t)(gi_fmt ? (a > 0 ? sub + bi : GI_BLACK) : (a > 0 ? bi : 0)) << (8 * k);
    }
    uint8_t *out = to_image ? im.map : maps + (size_t)e * NPIX;
    reinterpret_cast<uint32_t *>(out)[q] = packed;
}

// ------------------------------------------------------------------------------------------------
// k_pyramid: as_rgba() (lib.rs:550-577) + Rgb->LinearRgb + the 6-scale pyramid of ssimulacra2
// (downscale_by_2 on linear RGB, linear_rgb_to_xyb, make_positive_xyb) for one rendered candidate
// (SRC=false: every scale-0 pixel is a table lookup of its palette entry) or for the source image
// (SRC=true: lib.rs:506-516, alpha ignored).  One CTA walks 32x32 regions; a region yields 16x16,
// 8x8, 4x4, 2x2 and 1x1 pixels of the coarser scales entirely in shared memory.  Planes are written
// twice: row-major [ch][y][x] for the vertical pass and the maps, column-major [ch][x][y] for the
// horizontal pass (whose lanes walk along x, one lane per row).
// grid = (16, E), block 256.
// ------------------------------------------------------------------------------------------------
template <bool SRC>
__global__ void __launch_bounds__(256) k_pyramid(const ImgDev *imgs, const CandEntry *cents, int ncand, int e0, int S,
                                                 int CS, int has_ovr, const uint8_t *maps, int from_image,
                                                 float *xyb_rm_base, int gi_fmt) {
    // SRC: every scale in both layouts (the column-major copy feeds k_blur_h); candidates: only the row-major planes
    // of scales >= 1 (the scorer renders scale 0 itself from the palette_map)
    constexpr bool lean = !SRC;
    __shared__ float s_lin[SRC ? 1 : MAX_ENTRIES + 1][3];
    __shared__ float s_xyb[SRC ? 1 : MAX_ENTRIES + 1][3];
    __shared__ float tile[3][32][33];
    __shared__ float linbuf[2][3][256];
    const int e = blockIdx.y, ea = e0 + e, img = SRC ? e : ea / ncand, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const ImgDev im = imgs[img];
    float *rm = SRC ? const_cast<float *>(im.xyb_rm) : xyb_rm_base + (size_t)e * EVAL_XYB_FLOATS;
    float *cm = SRC ? const_cast<float *>(im.xyb_cm) : nullptr;
    // from_image: 1 = the image's own palette_map, 2 = its prepared base assignment (gi format, assign_delta.cuh)
    const uint8_t *map = SRC ? nullptr : (from_image == 2 ? im.base_gi : (from_image ? im.map : maps + (size_t)e * NPIX));
    if (!SRC) {
        const int ovr = has_ovr >= 0 ? cents[ea].slot : -1;   // the entry this evaluation replaces
        for (int j = tid; j < CS; j += 256) {
            const bool o = (j == ovr);
            for (int c = 0; c < 3; c++) {
                s_lin[j][c] = o ? cents[ea].lin[c] : im.tables->lin[j][c];
                s_xyb[j][c] = o ? cents[ea].xyb[c] : im.tables->xyb[j][c];
            }
        }
        if (tid < 3) {
            s_lin[BLACK][tid] = im.tables->lin[BLACK][tid];
            s_xyb[BLACK][tid] = im.tables->xyb[BLACK][tid];
        }
    }
    for (int reg = blockIdx.x; reg < 64; reg += gridDim.x) {
        __syncthreads();
        const int bx = reg & 7, by = reg >> 3;
        const int qx = tid & 15, qy = tid >> 4;
        float lin[4][3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int dx = k & 1, dy = k >> 1;
            const int x = bx * 32 + 2 * qx + dx, y = by * 32 + 2 * qy + dy;
            uchar4 p = make_uchar4(0, 0, 0, 0);
            if (SRC || !gi_fmt) p = __ldg(im.rgba + y * W + x);
            float xv, yv, bv;
            if (SRC) {
                lin[k][0] = c_lin_lut[p.x];
                lin[k][1] = c_lin_lut[p.y];
                lin[k][2] = c_lin_lut[p.z];
                lin_to_pxyb(lin[k][0], lin[k][1], lin[k][2], xv, yv, bv);
            } else {
                int gi;
                if (gi_fmt) {
                    gi = map[y * W + x];
                    gi = gi == GI_BLACK ? BLACK : gi;
                } else {
                    gi = p.w > 0 ? im.tile_pal[(y >> 3) * 32 + (x >> 3)] * S + map[y * W + x] : BLACK;
                }
                lin[k][0] = s_lin[gi][0];
                lin[k][1] = s_lin[gi][1];
                lin[k][2] = s_lin[gi][2];
                xv = s_xyb[gi][0];
                yv = s_xyb[gi][1];
                bv = s_xyb[gi][2];
            }
            tile[0][2 * qy + dy][2 * qx + dx] = xv;
            tile[1][2 * qy + dy][2 * qx + dx] = yv;
            tile[2][2 * qy + dy][2 * qx + dx] = bv;
        }
        __syncthreads();
        // scale 0 out: 128-byte rows in both layouts
        if (!lean)
            for (int r = warp; r < 32; r += 8)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    rm[c * NPIX + (by * 32 + r) * W + bx * 32 + lane] = tile[c][r][lane];
                    cm[c * NPIX + (bx * 32 + r) * H + by * 32 + lane] = tile[c][lane][r];
                }
        // scales 1..5: m x m pixels of this region, m = 16, 8, 4, 2, 1
        float cur[3];
#pragma unroll
        for (int c = 0; c < 3; c++) cur[c] = (((lin[0][c] + lin[1][c]) + lin[2][c]) + lin[3][c]) * 0.25f;
        int m = 16;
        for (int L = 1; L < NSCALES; L++, m >>= 1) {
            const int d = W >> L;
            const bool active = tid < m * m;
            const int ax = tid % m, ay = tid / m;
            if (L > 1 && active) {
                const float(*prev)[256] = linbuf[L & 1];
                const int pm = 2 * m;
#pragma unroll
                for (int c = 0; c < 3; c++)
                    cur[c] = (((prev[c][(2 * ay) * pm + 2 * ax] + prev[c][(2 * ay) * pm + 2 * ax + 1]) +
                               prev[c][(2 * ay + 1) * pm + 2 * ax]) +
                              prev[c][(2 * ay + 1) * pm + 2 * ax + 1]) *
                             0.25f;
            }
            __syncthreads();  // everyone done reading tile / linbuf[L&1] of the previous level
            if (active) {
                float xv, yv, bv;
                lin_to_pxyb(cur[0], cur[1], cur[2], xv, yv, bv);
                tile[0][ay][ax] = xv;
                tile[1][ay][ax] = yv;
                tile[2][ay][ax] = bv;
                float(*nxt)[256] = linbuf[(L + 1) & 1];
#pragma unroll
                for (int c = 0; c < 3; c++) nxt[c][ay * m + ax] = cur[c];
                // the prepared base assignment also keeps its LINEAR scale-2 plane ([3][64][64]), in the scale-0 area of the buffer,
                // which candidates never use: k_assign_pyr reads it for the 4x4 blocks a candidate leaves unchanged
                if (!SRC && from_image == 2 && L == 2)
#pragma unroll
                    for (int c = 0; c < 3; c++) rm[c * 4096 + (by * m + ay) * 64 + bx * m + ax] = cur[c];
            }
            __syncthreads();
            if (active) {
                const size_t off = 3 * (size_t)scale_off(L);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    rm[off + (size_t)c * d * d + (by * m + ay) * d + bx * m + ax] = tile[c][ay][ax];
                    if (!lean) cm[off + (size_t)c * d * d + (bx * m + ay) * d + by * m + ax] = tile[c][ax][ay];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_blur_h / k_blur_v: the source side of SSIMULACRA2, once per image (the reference recomputes it in every
// error() call): mu1 = blur(i1) and s11 = blur(i1*i1) for every scale and channel with the recursive Gaussian of
// ssimulacra2 (sigma 1.5, radius 5, zero padding; 3 second-order sections per plane):
//   horizontal:  out_k = n2_k*(in[n-6] + in[n+4]) - prev2_k ; out_k = fma(-d1_k, prev_k, out_k)
//   vertical:    t = fma(prev_k, d1_k, prev2_k) ; out_k = fma(sum, n2_k, -t)
// k_blur_h: one lane per image row walking x = -4 .. d-1; inputs come from the column-major copy (consecutive lanes
// = consecutive rows = consecutive addresses), outputs are transposed through a per-warp shared-memory tile and
// stored as 128-byte row segments of the H planes h[scale][ch][plane][y][x].
// grid = ceil(rows / 128), block 128 (4 warps x 32 rows); dynamic smem = 4 * 2 * 32*33 floats.
// k_blur_v: one thread per image column walking y = -4 .. d-1 over the H planes, writes mu1 / s11.
// grid = ceil(columns / 128), block 128.  (The candidate side of both passes lives in k_score_v3.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_blur_h(int s, int nrows_total, const ImgDev *imgs, float *h_base) {
    constexpr int NPL = 2;
    extern __shared__ float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float(*tile)[32][33] = reinterpret_cast<float(*)[32][33]>(smem + warp * NPL * 32 * 33);
    const int d = W >> s, dd = d * d;
    const size_t off3 = 3 * (size_t)scale_off(s);
    const int row0 = (blockIdx.x * 4 + warp) * 32;
    if (row0 >= nrows_total) return;
    const int r = row0 + lane;
    const bool valid = r < nrows_total;
    const int rr = valid ? r : row0;
    const int y = rr % d, ch = (rr / d) % 3, e = rr / (3 * d);
    const float *in1 = imgs[e].xyb_cm + off3 + (size_t)ch * dd + y;
    const float n2_0 = c_n2[0], n2_1 = c_n2[1], n2_2 = c_n2[2];
    const float md1_0 = -c_d1[0], md1_1 = -c_d1[1], md1_2 = -c_d1[2];
    float prev[NPL][3], prev2[NPL][3];
#pragma unroll
    for (int p = 0; p < NPL; p++)
#pragma unroll
        for (int k = 0; k < 3; k++) prev[p][k] = prev2[p][k] = 0.0f;

    for (int n = -4; n < d; n++) {
        const int right = n + 4, left = n - 6;
        float b_r = 0.0f, b_l = 0.0f;
        if (right < d) b_r = __ldg(in1 + (size_t)right * d);
        if (left >= 0) b_l = __ldg(in1 + (size_t)left * d);
        const float sum[NPL] = {b_l + b_r, b_l * b_l + b_r * b_r};
#pragma unroll
        for (int p = 0; p < NPL; p++) {
            float o0 = sum[p] * n2_0 - prev2[p][0];
            float o1 = sum[p] * n2_1 - prev2[p][1];
            float o2 = sum[p] * n2_2 - prev2[p][2];
            o0 = __fmaf_rn(md1_0, prev[p][0], o0);
            o1 = __fmaf_rn(md1_1, prev[p][1], o1);
            o2 = __fmaf_rn(md1_2, prev[p][2], o2);
            prev2[p][0] = prev[p][0];
            prev2[p][1] = prev[p][1];
            prev2[p][2] = prev[p][2];
            prev[p][0] = o0;
            prev[p][1] = o1;
            prev[p][2] = o2;
            if (n >= 0) tile[p][lane][n & 31] = (o0 + o1) + o2;
        }
        if (n >= 0 && ((n & 31) == 31 || n == d - 1)) {
            __syncwarp();
            const int x0 = n & ~31, cnt = (n & 31) + 1;
            for (int q = 0; q < 32; q++) {
                const int r2 = row0 + q;
                if (r2 >= nrows_total) break;
                const int y2 = r2 % d, ch2 = (r2 / d) % 3, e2 = r2 / (3 * d);
                float *out = h_base + ((size_t)e2 * 3 * TOTPIX + off3) * NPL + (size_t)(ch2 * NPL) * dd + (size_t)y2 * d + x0;
                if (lane < cnt) {
#pragma unroll
                    for (int p = 0; p < NPL; p++) out[(size_t)p * dd + lane] = tile[p][q][lane];
                }
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(128) k_blur_v(int s, int ncols_total, const ImgDev *imgs, const float *h_base) {
    constexpr int NPL = 2;
    const int d = W >> s, dd = d * d;
    const size_t off3 = 3 * (size_t)scale_off(s);
    const int cidx = blockIdx.x * 128 + threadIdx.x;
    if (cidx >= ncols_total) return;
    const int x = cidx % d, ch = (cidx / d) % 3, e = cidx / (3 * d);
    const ImgDev im = imgs[e];
    const float *h = h_base + ((size_t)e * 3 * TOTPIX + off3) * NPL + (size_t)(ch * NPL) * dd + x;
    const size_t poff = off3 + (size_t)ch * dd + x;
    float *mu1p = im.mu1 + poff, *s11p = im.s11 + poff;
    const float n2_0 = c_n2[0], n2_1 = c_n2[1], n2_2 = c_n2[2];
    const float d1_0 = c_d1[0], d1_1 = c_d1[1], d1_2 = c_d1[2];
    float prev[NPL][3], prev2[NPL][3];
#pragma unroll
    for (int p = 0; p < NPL; p++)
#pragma unroll
        for (int k = 0; k < 3; k++) prev[p][k] = prev2[p][k] = 0.0f;
    for (int n = -4; n < d; n++) {
        const int bottom = n + 4, top = n - 6;
        float val[NPL];
#pragma unroll
        for (int p = 0; p < NPL; p++) {
            const float tv = top >= 0 ? __ldg(h + (size_t)p * dd + (size_t)top * d) : 0.0f;
            const float bv = bottom < d ? __ldg(h + (size_t)p * dd + (size_t)bottom * d) : 0.0f;
            const float sum = tv + bv;
            const float t0 = __fmaf_rn(prev[p][0], d1_0, prev2[p][0]);
            const float t1 = __fmaf_rn(prev[p][1], d1_1, prev2[p][1]);
            const float t2 = __fmaf_rn(prev[p][2], d1_2, prev2[p][2]);
            const float o0 = __fmaf_rn(sum, n2_0, -t0);
            const float o1 = __fmaf_rn(sum, n2_1, -t1);
            const float o2 = __fmaf_rn(sum, n2_2, -t2);
            prev2[p][0] = prev[p][0];
            prev2[p][1] = prev[p][1];
            prev2[p][2] = prev[p][2];
            prev[p][0] = o0;
            prev[p][1] = o1;
            prev[p][2] = o2;
            val[p] = (o0 + o1) + o2;
        }
        if (n < 0) continue;
        mu1p[(size_t)n * d] = val[0];
        s11p[(size_t)n * d] = val[1];
    }
}

__device__ __forceinline__ bool best_less(double ea, int ia, double eb, int ib) {
    return ea < eb || (ea == eb && ia < ib);
}

// ------------------------------------------------------------------------------------------------
// k_argmin: per image, strict-< first minimum over its candidates' errors (lib.rs:216, 258, 302),
// i.e. the lexicographic minimum of (error, candidate index).  grid = nimg, block 128.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_argmin(const double *scores, int ncand, int idx_base, Best *best) {
    __shared__ double s_e[4];
    __shared__ int s_i[4];
    const int img = blockIdx.x, tid = threadIdx.x;
    double be = __longlong_as_double(0x7ff0000000000000ll);  // +inf
    int bi = 0x7fffffff;
    for (int k = tid; k < ncand; k += 128) {
        const double v = scores[(size_t)img * ncand + k];
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
    if ((tid & 31) == 0) {
        s_e[tid >> 5] = be;
        s_i[tid >> 5] = bi;
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
        b.idx = bi == 0x7fffffff ? -1 : bi + idx_base;
        b.pad = 0;
        best[img] = b;
    }
}

// "no candidate was evaluated" records (a rank whose candidate slice is empty)
__global__ void k_no_best(Best *best, int nimg) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nimg) return;
    Best b;
    b.err = __longlong_as_double(0x7ff0000000000000ll);
    b.idx = -1;
    b.pad = 0;
    best[j] = b;
}

// k_merge_best: the cross-rank argmin after the all-gather: gathered[r][j] is rank r's (error, index)
// for image j, with indices already global; the lexicographic minimum reproduces the strict-< /
// lowest-index-wins rule of lib.rs:216 for any number of ranks.  One thread per image.
__global__ void k_merge_best(const Best *gathered, int nranks, int rank_stride /* records per rank, >= nimg */, int nimg, Best *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nimg) return;
    Best b = gathered[j];
    for (int r = 1; r < nranks; r++) {
        const Best o = gathered[(size_t)r * rank_stride + j];
        if (o.idx >= 0 && (b.idx < 0 || best_less(o.err, o.idx, b.err, b.idx))) b = o;
    }
    out[j] = b;
}

// ------------------------------------------------------------------------------------------------
// k_apply_best: the accept step of lib.rs:199/216-219/236 (random, channel: only if strictly better
// than the current error) or lib.rs:250/264/280 (NES: always the first minimum).  One thread per
// image; rewrites the palette entry and the image's cached error.
// ------------------------------------------------------------------------------------------------
// chosen[j] (optional) = index of the accepted candidate in the full list, or -1: what k_adopt_map copies the palette_map of.
__global__ void k_apply_best(const ImgDev *imgs, int nimg, int slot, const uint8_t *cand_all, int ncand_all,
                             const Best *best, int force, int *chosen = nullptr) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nimg) return;
    if (chosen) chosen[j] = -1;
    const Best b = best[j];
    if (b.idx < 0 || b.idx >= ncand_all) return;
    const ImgDev im = imgs[j];
    const uint8_t *c = cand_all + ((size_t)j * ncand_all + b.idx) * 3;
    if (c[0] > 32 || c[1] > 32 || c[2] > 32) return;   // never a SnesColor (flagged by k_tables): not applied
    if (force || b.err < *im.cur_err) {
        im.palette[3 * slot] = c[0];
        im.palette[3 * slot + 1] = c[1];
        im.palette[3 * slot + 2] = c[2];
        *im.cur_err = b.err;
        if (chosen) chosen[j] = b.idx;
    }
}

// k_apply_tile_move: the accept step for tile-reassignment candidates: image j takes its best move if that is strictly
// better than its current error (the rule of lib.rs:216-219 applied to tile_palettes instead of a palette entry).
__global__ void k_apply_tile_move(const ImgDev *imgs, int nimg, const TileMove *moves, int nmoves, const Best *best, uint8_t *applied) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nimg) return;
    const Best b = best[j];
    uint8_t took = 0;
    if (b.idx >= 0 && b.idx < nmoves) {
        const ImgDev im = imgs[j];
        if (b.err < *im.cur_err) {
            const TileMove m = moves[(size_t)j * nmoves + b.idx];
            im.tile_pal[m.tile] = (uint8_t)m.sub;
            *im.cur_err = b.err;
            took = 1;
        }
    }
    if (applied) applied[j] = took;
}

// The accept step of `nsteps` consecutive optimiser iterations of an image evaluated against ONE state (speculation: the
// reference runs them one after the other, lib.rs:889-933, and an iteration that finds nothing better leaves the state
// as it was, so the next iteration's evaluations against the old state are exactly the reference's).  best[j][s] is the
// first minimum of step s's candidates.  Steps are taken in order; the first whose best beats the image's current error
// (lib.rs:216-219) is applied and ends the run; a NES step (force, lib.rs:250) always takes its first minimum and ends the
// run only if that changes the entry's colour: consumed[j] = the last step's number + 1 (nsteps if none),
// chosen[j] = the accepted evaluation's index among the image's nsteps * ncand evaluations, or -1.
__global__ void k_apply_first_accept(const ImgDev *imgs, int nimg, const int *step_slot, int nsteps, const uint8_t *cand, int ncand,
                                     const Best *best, int force, int *consumed, int *chosen, double *err_before /* NES only, or null */) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nimg) return;
    const ImgDev im = imgs[j];
    int used = nsteps, pick = -1;
    for (int s = 0; s < nsteps; s++) {
        const Best b = best[(size_t)j * nsteps + s];
        if (b.idx < 0 || b.idx >= ncand) continue;
        const uint8_t *c = cand + (((size_t)j * nsteps + s) * ncand + b.idx) * 3;
        if (c[0] > 32 || c[1] > 32 || c[2] > 32) continue;
        const int slot = step_slot[s];
        const uint8_t *cur = im.palette + 3 * slot;
        if (force && c[0] == cur[0] && c[1] == cur[1] && c[2] == cur[2]) {   // NES step that leaves the colour as it is
            *im.cur_err = b.err;
            if (err_before) err_before[j] = b.err;
            continue;
        }
        if (force || b.err < *im.cur_err) {
            im.palette[3 * slot] = c[0];
            im.palette[3 * slot + 1] = c[1];
            im.palette[3 * slot + 2] = c[2];
            *im.cur_err = b.err;
            used = s + 1;
            pick = s * ncand + b.idx;
            break;
        }
    }
    consumed[j] = used;
    chosen[j] = pick;
}

// optimize() after an accepted candidate without running it again: the accepted evaluation's assignment is still in the
// scratch maps (gi format: tile_sub * S + index, GI_BLACK for a transparent pixel) and IS optimize() of the new state
// (lib.rs:236-237 recompute what lib.rs:212 computed for that candidate).  grid = (64, nimg), block 256, 4 pixels per thread.
__global__ void __launch_bounds__(256) k_adopt_map(const ImgDev *imgs, const int *chosen, const uint8_t *maps, int ncand_total, int S) {
    const int j = blockIdx.y, pick = chosen[j];
    if (pick < 0) return;
    const ImgDev im = imgs[j];
    const int q = blockIdx.x * 256 + threadIdx.x, px0 = q * 4, y = px0 >> 8, x = px0 & 255;
    const uint32_t g = reinterpret_cast<const uint32_t *>(maps + ((size_t)j * ncand_total + pick) * NPIX)[q];
    const int sub = im.tile_pal[(y >> 3) * 32 + (x >> 3)] * S;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int gi = (g >> (8 * k)) & 255;
        out |= (uint32_t)(gi == GI_BLACK ? 0 : gi - sub) << (8 * k);
    }
    reinterpret_cast<uint32_t *>(im.map)[q] = out;
}

// current error of each image := scores[j]  (after an error() pass over the images themselves)
__global__ void k_store_cur_err(const ImgDev *imgs, int nimg, const double *scores) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nimg) *imgs[j].cur_err = scores[j];
}

// out[j] := current error of image j
__global__ void k_load_cur_err(const ImgDev *imgs, int nimg, double *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nimg) out[j] = *imgs[j].cur_err;
}

// candidate lists built on the device: the 56 NES colours (lib.rs:252-253) or the 32 values of one
// channel of the current entry (lib.rs:296-297).  grid = nimg, block 64.
// With several steps per image (grid = (nimg, nsteps)) step s uses entry step_slot[s] and channel step_channel[s]; the lists
// of an image follow each other.
__global__ void k_make_cands(const ImgDev *imgs, int slot, int channel /* -1: NES */, uint8_t *cand, int ncand,
                             const int *step_slot = nullptr, const int *step_channel = nullptr) {
    const int j = blockIdx.x, k = threadIdx.x, s = blockIdx.y;
    if (k >= ncand) return;
    if (step_slot) {
        slot = step_slot[s];
        channel = step_channel[s];
    }
    uint8_t *o = cand + (((size_t)j * gridDim.y + s) * ncand + k) * 3;
    if (channel < 0) {
        o[0] = c_nes[k][0];
        o[1] = c_nes[k][1];
        o[2] = c_nes[k][2];
    } else {
        const uint8_t *cur = imgs[j].palette + 3 * slot;
        o[0] = channel == 0 ? (uint8_t)k : cur[0];
        o[1] = channel == 1 ? (uint8_t)k : cur[1];
        o[2] = channel == 2 ? (uint8_t)k : cur[2];
    }
}

// as_rgba() (lib.rs:550-577) of the image's current state.  grid = 256, block 256.
__global__ void k_as_rgba(ImgDev im, int S, uchar4 *out) {
    const int px = blockIdx.x * 256 + threadIdx.x;
    const int x = px & 255, y = px >> 8;
    const uchar4 p = im.rgba[px];
    uchar4 o = make_uchar4(0, 0, 0, 0);
    if (p.w > 0) {
        const uint8_t *c = im.palette + 3 * (im.tile_pal[(y >> 3) * 32 + (x >> 3)] * S + im.map[px]);
        o = snes_as_rgba(c[0], c[1], c[2]);
    }
    out[px] = o;
}

}  // namespace snes
