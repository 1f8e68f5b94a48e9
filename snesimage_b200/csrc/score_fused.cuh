// k_score_fused: the whole candidate side of SSIMULACRA2 for one (evaluation, channel) per CTA, with
// every intermediate plane on chip.
//
//   error() = 100 - compute_frame_ssimulacra2(src, dst)      (lib.rs:503-548, ssimulacra2 0.5.1)
//
// Per scale the crate blurs {i2, i2*i2, i1*i2} with a separable recursive Gaussian (horizontal pass
// along rows, then vertical pass along columns; both are 3 second-order IIR sections, so each line is
// a serial chain) and feeds mu2, s22, s12 together with the image's precomputed mu1, s11, i1 into
// ssim_map and edge_diff_map.  The two passes want opposite thread mappings (thread = row, then
// thread = column), which is a transposition; here it happens in shared memory, tile by tile:
//
//   for column block j (BW columns)            <- horizontal IIR state stays in registers across j
//     for row half h (HB <= 128 rows)          <- vertical IIR state stays in registers across h
//       stage   i2 and i1*i2 of the tile (+6/+4 halo columns) into smem, coalesced; at scale 0 the
//               rendered pixel is a table lookup of its palette entry (as_rgba, lib.rs:550-577)
//       H pass  thread = row, walks BW columns, writes 3 blurred-H planes to smem (pitch BW+1)
//       V pass  thread = (column, plane), walks the rows; output n needs H rows n+4 and n-6, so it
//               lags 4 rows and keeps the last 10 H rows of the previous half; results overwrite
//               the H rows they have just consumed
//       maps    all threads: ssim_map + edge_diff_map terms of the tile in f64, accumulated per thread
//   block-reduce the six f64 sums of the (scale, channel) in a fixed order -> partials
//
// Bit-exactness: the f32 operation order of both passes and of the maps is the oracle's (and the
// previous multi-kernel path's), so every f32 intermediate is identical; only the order of the f64
// sums differs.  HBM traffic per evaluation is the palette_map, the image's source planes (shared by
// all candidates of the image -> L2) and the small coarse-scale XYB pyramid; no blur plane leaves the SM.
#pragma once
#include "kernels.cuh"

namespace snes {

constexpr int FUSED_THREADS = 256;
constexpr int FUSED_WARPS = FUSED_THREADS / 32;

template <int BW>
struct FusedSmem {
    static constexpr int HB = 128;
    static constexpr int NCOL = BW + 12;  // staged columns c0-8 .. c0+BW+3 (taps reach c0-6 .. c0+BW+3), 16-byte chunks
    static constexpr int IP = NCOL;       // pitch = 44 floats at BW = 32: lane = row reads are skewed to stay conflict-free
    static constexpr int NCH = NCOL / 4;
    static constexpr int HP = BW + 1;
    alignas(16) float in2[HB + 4][IP];   // i2 of rows r0-4 .. r0+HB-1 (the 4 extra rows serve the lagging maps)
    alignas(16) float in1[HB + 4][IP];   // i1 of the same tile
    float hout[3][HB + 10][HP];
    float xyb[MAX_ENTRIES + 1];
    double red[FUSED_WARPS][NSUMS];
    uint32_t raw[HB + 4][NCH];           // scale 0: palette_map bytes of the tile as aligned 4-pixel words
    uint8_t tp[NTILES];
};

struct FusedArgs {
    const ImgDev *imgs;
    const CandEntry *cents;
    int ncand, e0, S, CS, ovr;
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

// x / y for 1 <= y < 2^100 in f64 without the slow-path checks of the generic division: f32 reciprocal seed,
// two Newton steps, one residual correction (Markstein).  Correctly rounded except for rare last-bit cases.
__device__ __forceinline__ double div64_fast(double x, double y) {
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)y));   // MUFU.RCP seed, ~2^-23
    double r = (double)rf;
    double e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    const double q = x * r;
    return fma(fma(-y, q, x), r, q);
}

template <int D, int BW0, typename SM>
__device__ __forceinline__ void fused_scale(SM &sm, const FusedArgs &a, const ImgDev &im, const uint8_t *map, int e, int ea,
                                            int ch, int scale) {
    constexpr int BW = D < BW0 ? D : BW0;     // column block width at this scale
    constexpr int HB = D < 128 ? D : 128;     // rows per half
    constexpr int NH = D / HB;
    constexpr int NJ = D / BW;
    constexpr int RPW = 32 / BW;              // rows one warp covers per maps iteration
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t poff = 3 * (size_t)scale_off(scale) + (size_t)ch * D * D;
    const float *i1p = im.xyb_rm + poff;
    const float *mu1p = im.mu1 + poff;
    const float *s11p = im.s11 + poff;
    const float *i2p = a.xyb_rm + (size_t)e * EVAL_XYB_FLOATS + poff;  // unused at scale 0
    const float n2_0 = c_n2[0], n2_1 = c_n2[1], n2_2 = c_n2[2];
    const float d1_0 = c_d1[0], d1_1 = c_d1[1], d1_2 = c_d1[2];
    const float md1_0 = -d1_0, md1_1 = -d1_1, md1_2 = -d1_2;

    float hp[NH][2][3], hq[NH][2][3];  // horizontal IIR state of this thread's row(s) and plane slot(s): prev, prev2
#pragma unroll
    for (int h = 0; h < NH; h++)
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int k = 0; k < 3; k++) hp[h][p][k] = hq[h][p][k] = 0.0f;
    double acc[NSUMS] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};

    for (int j = 0; j < NJ; j++) {
        const int c0 = j * BW;
        float vp[3] = {0.0f, 0.0f, 0.0f}, vq[3] = {0.0f, 0.0f, 0.0f};  // vertical IIR state of (column, plane)
#pragma unroll
        for (int h = 0; h < NH; h++) {
            const int r0 = h * HB;
            const int y_lo = r0 - 4 < 0 ? 0 : r0 - 4;        // first staged image row; buffer row = y - (r0 - 4)
            const int nrows = r0 + HB - y_lo;
            // ---- stage the tile (rows y_lo .. r0+HB-1, columns c0-8 .. c0+BW+3; zero outside the image) ----------
            // Everything goes global -> smem with cp.async (no registers, all of a thread's loads in flight at once):
            // i1 (and i2 at scales >= 1) as 16-byte chunks, at scale 0 the palette_map bytes as aligned words that the
            // fetching thread converts after the wait -- the rendered pixel is a table lookup of its palette entry
            // (as_rgba, lib.rs:550-577).  A warp iteration covers two rows: lane = (row parity, chunk).
            {
                const int ry_lo = y_lo - (r0 - 4);
                const int w4 = lane & 15, rs = lane >> 4;
                const int x0 = c0 - 8 + 4 * w4;
                if (w4 < SM::NCH) {
                    const bool inside = x0 >= 0 && x0 < D;   // D and x0 are multiples of 4: a chunk is all in or all out
                    const int rr0 = 2 * warp + rs;
                    const float *g1 = i1p + (size_t)(y_lo + rr0) * D + x0;
                    const float *g2 = i2p + (size_t)(y_lo + rr0) * D + x0;
                    const uint8_t *gm = map + (y_lo + rr0) * W + x0;
                    unsigned so1 = smem_addr(&sm.in1[ry_lo + rr0][4 * w4]);
                    unsigned so2 = smem_addr(&sm.in2[ry_lo + rr0][4 * w4]);
                    unsigned sor = smem_addr(&sm.raw[ry_lo + rr0][w4]);
                    for (int r = rr0; r < nrows; r += 2 * FUSED_WARPS) {
                        if (inside) {
                            cp_async16(so1, g1);
                            if (D != W) cp_async16(so2, g2);
                            else cp_async4(sor, gm);
                        } else {
                            *reinterpret_cast<float4 *>(&sm.in1[ry_lo + r][4 * w4]) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            *reinterpret_cast<float4 *>(&sm.in2[ry_lo + r][4 * w4]) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        }
                        g1 += 2 * FUSED_WARPS * D;
                        g2 += 2 * FUSED_WARPS * D;
                        gm += 2 * FUSED_WARPS * W;
                        so1 += 2 * FUSED_WARPS * SM::IP * 4;
                        so2 += 2 * FUSED_WARPS * SM::IP * 4;
                        sor += 2 * FUSED_WARPS * SM::NCH * 4;
                    }
                    if (D == W && inside) {
                        cp_async_wait_all();
                        for (int r = rr0; r < nrows; r += 2 * FUSED_WARPS) {
                            const int y = y_lo + r, ry = ry_lo + r;
                            const uint32_t mw = sm.raw[ry][w4];
                            uint32_t aw = 0xffffffffu;
                            int sub = 0;
                            if (!a.gi_fmt) {   // palette_map format (error() of the image's own state): needs tile and alpha
                                aw = __ldg(reinterpret_cast<const uint32_t *>(im.alpha + y * W + x0));
                                sub = sm.tp[(y >> 3) * 32 + (x0 >> 3)];
                            }
                            float4 v;
                            v.x = sm.xyb[(aw & 255u) ? sub + (mw & 255u) : BLACK];
                            v.y = sm.xyb[((aw >> 8) & 255u) ? sub + ((mw >> 8) & 255u) : BLACK];
                            v.z = sm.xyb[((aw >> 16) & 255u) ? sub + ((mw >> 16) & 255u) : BLACK];
                            v.w = sm.xyb[(aw >> 24) ? sub + (mw >> 24) : BLACK];
                            *reinterpret_cast<float4 *>(&sm.in2[ry][4 * w4]) = v;
                        }
                    }
                }
            }
            cp_async_wait_all();
            __syncthreads();
            // ---- horizontal pass: thread = (row r0 + (t & 127), plane group t >> 7) -------------------------------------
            // group 0 (warps 0-3) runs the planes {i2, i2*i2}, group 1 (warps 4-7) the plane {i1*i2}: all 8 warps busy.
            // The 16-byte staging forces an even row pitch, so lanes (= rows) are skewed by (row >> 3) & 3 columns:
            // bank = 44*row + column is then distinct across the warp.  Tap columns in the tile: n-6 -> q+2, n+4 -> q+12.
            if ((t & 127) < HB) {
                const int row = t & 127, skew = (row >> 3) & 3;
                const float *r2 = sm.in2[row + 4], *r1 = sm.in1[row + 4];
                auto section = [&](float sum, int p) {   // p = plane slot of this thread's state
                    float o0 = sum * n2_0 - hq[h][p][0];
                    float o1 = sum * n2_1 - hq[h][p][1];
                    float o2 = sum * n2_2 - hq[h][p][2];
                    o0 = __fmaf_rn(md1_0, hp[h][p][0], o0);
                    o1 = __fmaf_rn(md1_1, hp[h][p][1], o1);
                    o2 = __fmaf_rn(md1_2, hp[h][p][2], o2);
                    hq[h][p][0] = hp[h][p][0];
                    hq[h][p][1] = hp[h][p][1];
                    hq[h][p][2] = hp[h][p][2];
                    hp[h][p][0] = o0;
                    hp[h][p][1] = o1;
                    hp[h][p][2] = o2;
                    return (o0 + o1) + o2;
                };
                if (t < 128) {
                    if (j == 0) {  // warm-up n = -4..-1: left taps are outside the image
#pragma unroll
                        for (int n = -4; n < 0; n++) {
                            const float ar = r2[n + 12];
                            section(ar, 0);
                            section(ar * ar, 1);
                        }
                    }
#pragma unroll 4
                    for (int qq = 0; qq < BW + 3; qq++) {  // n = c0 + q
                        const int q = qq - skew;
                        if (q >= 0 && q < BW) {
                            const float ar = r2[q + 12], al = r2[q + 2];
                            sm.hout[0][10 + row][q] = section(al + ar, 0);
                            sm.hout[1][10 + row][q] = section(al * al + ar * ar, 1);
                        }
                    }
                } else {
                    if (j == 0) {
#pragma unroll
                        for (int n = -4; n < 0; n++) section(r1[n + 12] * r2[n + 12], 0);
                    }
#pragma unroll 4
                    for (int qq = 0; qq < BW + 3; qq++) {
                        const int q = qq - skew;
                        if (q >= 0 && q < BW) {
                            const float ar = r2[q + 12], al = r2[q + 2], br = r1[q + 12], bl = r1[q + 2];
                            sm.hout[2][10 + row][q] = section(bl * al + br * ar, 0);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- vertical pass: thread = (column, plane); buffer row index of image row g is g - r0 + 10 ----------
            if (t < 3 * BW) {
                const int col = t % BW, p = t / BW;
                float *hb = &sm.hout[p][0][col];   // row stride SM::HP
                auto vstep = [&](float sum, float *store) {
                    const float t0 = __fmaf_rn(vp[0], d1_0, vq[0]);
                    const float t1 = __fmaf_rn(vp[1], d1_1, vq[1]);
                    const float t2 = __fmaf_rn(vp[2], d1_2, vq[2]);
                    const float o0 = __fmaf_rn(sum, n2_0, -t0);
                    const float o1 = __fmaf_rn(sum, n2_1, -t1);
                    const float o2 = __fmaf_rn(sum, n2_2, -t2);
                    vq[0] = vp[0];
                    vq[1] = vp[1];
                    vq[2] = vp[2];
                    vp[0] = o0;
                    vp[1] = o1;
                    vp[2] = o2;
                    if (store) *store = (o0 + o1) + o2;   // slot of H row n-6, consumed by this step
                };
                // buffer indices for output n: top tap (n-6) -> n - r0 + 4, bottom tap (n+4) -> n - r0 + 14, store -> n - r0 + 4
                int n = r0 - 4;
                const int n_end = (h == NH - 1) ? D : r0 + HB - 4;      // exclusive
                const int n_main_end = (h == NH - 1) ? D - 4 : n_end;   // bottom tap inside the image below this
                if (h == 0) {
                    for (; n < 0; n++) vstep(hb[(n + 14) * SM::HP], nullptr);                      // warm-up
                    for (; n < 6 && n < n_end; n++)                                                   // no top tap yet
                        vstep(n < D - 4 ? hb[(n + 14) * SM::HP] : 0.0f, &hb[(n + 4) * SM::HP]);
                }
                const float *pt = hb + (n - r0 + 4) * SM::HP;
#pragma unroll 4
                for (; n < n_main_end; n++, pt += SM::HP) vstep(pt[0] + pt[10 * SM::HP], const_cast<float *>(pt));
                for (; n < n_end; n++, pt += SM::HP) vstep(pt[0] + 0.0f, const_cast<float *>(pt));     // bottom tap below the image
            }
            __syncthreads();
            // ---- ssim_map + edge_diff_map over the rows the vertical pass has just finished ---------------------------
            // warp iteration = RPW rows x BW columns; global loads of UM iterations issued ahead of the f64 math
            {
                constexpr int UM = 4;
                const int n_begin = r0 - 4 < 0 ? 0 : r0 - 4;
                const int n_end = (h == NH - 1) ? D : r0 + HB - 4;
                const int col = lane % BW, rsub = lane / BW;
                const float *mu1c = mu1p + c0 + col, *s11c = s11p + c0 + col;
#pragma unroll 1
                for (int nb = n_begin + warp * RPW + rsub; nb < n_end; nb += UM * FUSED_WARPS * RPW) {
                    float mu1v[UM], s11v[UM];
#pragma unroll
                    for (int u = 0; u < UM; u++) {
                        const int n = nb + u * FUSED_WARPS * RPW;
                        mu1v[u] = s11v[u] = 0.0f;
                        if (n < n_end) {
                            mu1v[u] = __ldg(mu1c + n * D);
                            s11v[u] = __ldg(s11c + n * D);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UM; u++) {
                        const int n = nb + u * FUSED_WARPS * RPW;
                        if (n < n_end) {
                            const int bi = n - r0 + 4;
                            const float mu2 = sm.hout[0][bi][col], s22 = sm.hout[1][bi][col], s12 = sm.hout[2][bi][col];
                            const float i1 = sm.in1[bi][col + 8], i2 = sm.in2[bi][col + 8];
                            const float mu1 = mu1v[u], s11 = s11v[u];
                            const float mu11 = mu1 * mu1, mu22 = mu2 * mu2, mu12 = mu1 * mu2;
                            const float mu_diff = mu1 - mu2;
                            const float num_m = __fmaf_rn(mu_diff, -mu_diff, 1.0f);
                            const float num_s = __fmaf_rn(2.0f, s12 - mu12, 0.0009f);
                            const float denom_s = (s11 - mu11) + (s22 - mu22) + 0.0009f;
                            // d = max(1 - q, 0): positive iff q < 1 (1 - q is exact in sign; NaN contributes 0 like fmax)
                            const float qf = (num_m * num_s) / denom_s;
                            if (qf < 1.0f) {
                                const double dv = 1.0 - (double)qf;
                                acc[0] += dv;
                                const double dv2 = dv * dv;
                                acc[1] += dv2 * dv2;
                            }
                            // d1 = (1 + |i2 - mu2|) / (1 + |i1 - mu1|) - 1; artifact = max(d1, 0), detail_lost = max(-d1, 0):
                            // exactly one of them is |d1|, selected by the sign bit
                            const double d1v = div64_fast(1.0 + (double)fabsf(i2 - mu2), 1.0 + (double)fabsf(i1 - mu1)) - 1.0;
                            const double ad = fabs(d1v);
                            const double ad2 = ad * ad;
                            const double ad4 = ad2 * ad2;
                            if (__double2hiint(d1v) >= 0) {
                                acc[2] += ad;
                                acc[3] += ad4;
                            } else {
                                acc[4] += ad;
                                acc[5] += ad4;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // ---- keep the last 10 H rows of this half for the next one
            if (h + 1 < NH) {
                for (int idx = t; idx < 3 * 10 * BW; idx += FUSED_THREADS) {
                    const int p = idx / (10 * BW), rem = idx - p * 10 * BW, rr = rem / BW, col = rem - rr * BW;
                    sm.hout[p][rr][col] = sm.hout[p][HB + rr][col];
                }
                __syncthreads();
            }
        }
    }
    // ---- fixed-order block reduction of the six sums of this (scale, channel)
#pragma unroll
    for (int q = 0; q < NSUMS; q++) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm.red[warp][q] = v;
    }
    __syncthreads();
    if (t < NSUMS) {
        double v = sm.red[0][t];
        for (int w2 = 1; w2 < FUSED_WARPS; w2++) v += sm.red[w2][t];
        a.partials[(size_t)ea * (NSCALES * 3 * NSUMS) + ((size_t)scale * 3 + ch) * NSUMS + t] = v;
    }
    __syncthreads();
}

// grid = (3 channels, evaluations of the chunk), block = FUSED_THREADS, dynamic smem = sizeof(FusedSmem<BW>)
template <int BW>
__global__ void __launch_bounds__(FUSED_THREADS, BW == 16 ? 3 : 2) k_score_fused(const FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem<BW> &sm = *reinterpret_cast<FusedSmem<BW> *>(smem_raw);
    const int ch = blockIdx.x, e = blockIdx.y, ea = a.e0 + e, img = ea / a.ncand, t = threadIdx.x;
    const ImgDev im = a.imgs[img];
    const uint8_t *map = a.from_image ? im.map : a.maps + (size_t)e * NPIX;
    for (int i = t; i < a.CS; i += FUSED_THREADS) sm.xyb[i] = (i == a.ovr) ? a.cents[ea].xyb[ch] : im.tables->xyb[i][ch];
    if (t == 0) {
        sm.xyb[BLACK] = im.tables->xyb[BLACK][ch];
        if (a.gi_fmt) sm.xyb[GI_BLACK] = im.tables->xyb[BLACK][ch];   // C*S <= 255 there: slot 255 is free
    }
    for (int i = t; i < NTILES; i += FUSED_THREADS) sm.tp[i] = (uint8_t)(im.tile_pal[i] * a.S);  // subpalette offset, <= 255
    __syncthreads();
    fused_scale<256, BW>(sm, a, im, map, e, ea, ch, 0);
    fused_scale<128, BW>(sm, a, im, map, e, ea, ch, 1);
    fused_scale<64, BW>(sm, a, im, map, e, ea, ch, 2);
    fused_scale<32, BW>(sm, a, im, map, e, ea, ch, 3);
    fused_scale<16, BW>(sm, a, im, map, e, ea, ch, 4);
    fused_scale<8, BW>(sm, a, im, map, e, ea, ch, 5);
}

// k_pool for the fused path's partials layout [E][scale][channel][6]
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
