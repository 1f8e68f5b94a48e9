// CIELAB side of libsnesgpu (sm_100a): palette 0.7.6's Srgb<u8> -> Lab<D65,f32>, Lab<f64> -> Srgb<u8>
// and Ciede2000, as used by color_distance_cielab (lib.rs:1090-1100), new_nes_only (lib.rs:640-660)
// and the Lab branches of initialize_tiles / recalculate_palette (lib.rs:101-106, 141-142, 344-346,
// 369-371).  The reference memoises (RGBA8, RGBA8) -> distance in an unbounded global cache; here
// the palette side is a 32,768-entry BGR555 -> Lab table in global memory (L2-resident, 512 KiB)
// and the target side is the image's per-pixel Lab plane, or computed on the fly when dithering.
//
// f32 operations are the IEEE operations in the oracle's order (-fmad=false).  Transcendentals are
// CUDA's (cbrt via f64, atan2f/sinf/cosf/expf): results agree with the glibc-based oracle to a few
// ulp, which is why Lab-mode parity is stated with a tolerance (tests/test_lab_gpu.py).
#pragma once
#include "common.cuh"

namespace snes {

__constant__ float c_srgb_lin_lut[256];  // palette Srgb::into_linear(v/255), host-built with libm powf

__device__ __forceinline__ float lab_cbrtf(float x) { return (float)cbrt((double)x); }

__device__ __forceinline__ void lin_to_lab(float c0, float c1, float c2, float &L, float &A, float &B) {
    float x = ((float)0.4124564 * c0 + (float)0.3575761 * c1) + (float)0.1804375 * c2;
    float y = ((float)0.2126729 * c0 + (float)0.7151522 * c1) + (float)0.0721750 * c2;
    float z = ((float)0.0193339 * c0 + (float)0.1191920 * c1) + (float)0.9503041 * c2;
    x /= (float)0.95047;
    y /= 1.0f;
    z /= (float)1.08883;
    const float eps = (float)((6.0 / 29.0) * (6.0 / 29.0) * (6.0 / 29.0));
    const float kappa = (float)(841.0 / 108.0);
    const float delta = (float)(4.0 / 29.0);
    const float fx = x > eps ? lab_cbrtf(x) : kappa * x + delta;
    const float fy = y > eps ? lab_cbrtf(y) : kappa * y + delta;
    const float fz = z > eps ? lab_cbrtf(z) : kappa * z + delta;
    L = fy * 116.0f - 16.0f;
    A = (fx - fy) * 500.0f;
    B = (fy - fz) * 200.0f;
}

__device__ __forceinline__ void srgb8_to_lab(int r, int g, int b, float &L, float &A, float &B) {
    lin_to_lab(c_srgb_lin_lut[r], c_srgb_lin_lut[g], c_srgb_lin_lut[b], L, A, B);
}

// palette IntoStimulus<u8> for f64: clamp to [0,1], x255, round half to even.
__device__ __forceinline__ uint8_t f64_to_u8_stimulus(double v) {
    double s = v * 255.0;
    if (!(s > 0.0)) return 0;  // also NaN
    if (s > 255.0) s = 255.0;
    return (uint8_t)rint(s);
}

// Lab<D65,f64> -> Srgb<u8> (lib.rs:141-142, 369-371)
__device__ inline void lab_f64_to_srgb8(const double lab[3], uint8_t out[3]) {
    const double fy = (lab[0] + 16.0) / 116.0;
    const double fx = fy + lab[1] / 500.0;
    const double fz = fy - lab[2] / 200.0;
    const double eps = 6.0 / 29.0, kappa = 108.0 / 841.0, delta = 4.0 / 29.0;
    const double x = (fx > eps ? fx * fx * fx : (fx - delta) * kappa) * 0.95047;
    const double y = (fy > eps ? fy * fy * fy : (fy - delta) * kappa) * 1.0;
    const double z = (fz > eps ? fz * fz * fz : (fz - delta) * kappa) * 1.08883;
    double lin[3];
    lin[0] = (3.2404542 * x + -1.5371385 * y) + -0.4985314 * z;
    lin[1] = (-0.9692660 * x + 1.8760108 * y) + 0.0415560 * z;
    lin[2] = (0.0556434 * x + -0.2040259 * y) + 1.0572252 * z;
    for (int i = 0; i < 3; i++) {
        const double v = lin[i] <= 0.0031308 ? 12.92 * lin[i] : 1.055 * pow(lin[i], 1.0 / 2.4) - 0.055;
        out[i] = f64_to_u8_stimulus(v);
    }
}

// palette::color_difference::Ciede2000 (Sharma/Wu/Dalal 2005, kL=kC=kH=1) in f32.
// Argument order as the reference calls it: lab1 = palette colour, lab2 = target (lib.rs:783).
__device__ inline float ciede2000(float l1, float a1, float b1, float l2, float a2, float b2) {
    const float pi_over_180 = (float)(3.14159265358979323846 / 180.0);
    const float p25_7 = 6103515625.0f;
    const float c1 = sqrtf(a1 * a1 + b1 * b1);
    const float c2 = sqrtf(a2 * a2 + b2 * b2);
    const float delta_l_prime = l2 - l1;
    const float l_bar = (l1 + l2) / 2.0f;
    const float c_bar = (c1 + c2) / 2.0f;
    const float c_bar2 = c_bar * c_bar;
    const float c_bar7 = c_bar2 * c_bar2 * c_bar2 * c_bar;
    const float g = 0.5f * (1.0f - sqrtf(c_bar7 / (c_bar7 + p25_7)));
    const float a1p = a1 * (1.0f + g);
    const float a2p = a2 * (1.0f + g);
    const float c1p = sqrtf(a1p * a1p + b1 * b1);
    const float c2p = sqrtf(a2p * a2p + b2 * b2);
    float h1p = 0.0f, h2p = 0.0f;
    if (!(b1 == 0.0f && a1p == 0.0f)) {
        h1p = atan2f(b1, a1p) / pi_over_180;
        if (h1p < 0.0f) h1p += 360.0f;
    }
    if (!(b2 == 0.0f && a2p == 0.0f)) {
        h2p = atan2f(b2, a2p) / pi_over_180;
        if (h2p < 0.0f) h2p += 360.0f;
    }
    const float h_diff = h2p - h1p;
    const float h_abs = fabsf(h_diff);
    const bool zero_chroma = (c1p == 0.0f || c2p == 0.0f);
    float delta_h_prime;
    if (zero_chroma) delta_h_prime = 0.0f;
    else if (h_abs <= 180.0f) delta_h_prime = h_diff;
    else if (h2p <= h1p) delta_h_prime = h_diff + 360.0f;
    else delta_h_prime = h_diff - 360.0f;
    const float delta_big_h = 2.0f * sqrtf(c1p * c2p) * sinf(delta_h_prime / 2.0f * pi_over_180);
    float h_bar;
    if (zero_chroma) h_bar = h1p + h2p;
    else if (h_abs > 180.0f) {
        if (h1p + h2p < 360.0f) h_bar = (h1p + h2p + 360.0f) / 2.0f;
        else h_bar = (h1p + h2p - 360.0f) / 2.0f;
    } else h_bar = (h1p + h2p) / 2.0f;
    const float lb50 = (l_bar - 50.0f) * (l_bar - 50.0f);
    const float c_bar_p = (c1p + c2p) / 2.0f;
    const float t = 1.0f - 0.17f * cosf((h_bar - 30.0f) * pi_over_180) + 0.24f * cosf((2.0f * h_bar) * pi_over_180) +
                    0.32f * cosf((3.0f * h_bar + 6.0f) * pi_over_180) - 0.20f * cosf((4.0f * h_bar - 63.0f) * pi_over_180);
    const float s_l = 1.0f + (0.015f * lb50) / sqrtf(20.0f + lb50);
    const float s_c = 1.0f + 0.045f * c_bar_p;
    const float s_h = 1.0f + 0.015f * c_bar_p * t;
    const float hb = (h_bar - 275.0f) / 25.0f;
    const float delta_theta = 30.0f * expf(-(hb * hb));
    const float cbp2 = c_bar_p * c_bar_p;
    const float cbp7 = cbp2 * cbp2 * cbp2 * c_bar_p;
    const float r_c = 2.0f * sqrtf(cbp7 / (cbp7 + p25_7));
    const float r_t = -r_c * sinf(2.0f * delta_theta * pi_over_180);
    const float delta_c_prime = c2p - c1p;
    const float tl = delta_l_prime / s_l;
    const float tc = delta_c_prime / s_c;
    const float th = delta_big_h / s_h;
    return sqrtf(tl * tl + tc * tc + th * th + r_t * tc * th);
}

// SnesColor 5-bit value -> index into the BGR555 table.  round(v/8) can store 32 (lib.rs:396-400),
// whose as_rgba() wraps to 8 == as_rgba(1) (lib.rs:664).  Values above 32 are refused at the boundary (host lists) or
// flagged by k_tables (device lists); the mask keeps the table lookup in range whatever arrives.
__device__ __forceinline__ int bgr555_index(int r5, int g5, int b5) {
    r5 = r5 == 32 ? 1 : (r5 & 31);
    g5 = g5 == 32 ? 1 : (g5 & 31);
    b5 = b5 == 32 ? 1 : (b5 & 31);
    return r5 | (g5 << 5) | (b5 << 10);  // SnesColor::as_u16, lib.rs:679-681
}

// k_build_lab_table: Lab of as_rgba(entry) for all 32,768 BGR555 words.  grid 128 x block 256.
__global__ void __launch_bounds__(256) k_build_lab_table(float4 *table) {
    const int w = blockIdx.x * 256 + threadIdx.x;
    const uchar4 c = snes_as_rgba(w & 31, (w >> 5) & 31, (w >> 10) & 31);
    float L, A, B;
    srgb8_to_lab(c.x, c.y, c.z, L, A, B);
    table[w] = make_float4(L, A, B, 0.0f);
}

// k_image_lab: Lab<D65,f32> of every original pixel (alpha ignored).  grid 256 x block 256.
__global__ void __launch_bounds__(256) k_image_lab(const uchar4 *rgba, float4 *lab) {
    const int px = blockIdx.x * 256 + threadIdx.x;
    const uchar4 p = rgba[px];
    float L, A, B;
    srgb8_to_lab(p.x, p.y, p.z, L, A, B);
    lab[px] = make_float4(L, A, B, 0.0f);
}

// k_assign_lab: optimize() (lib.rs:425-501) without dithering, CIEDE2000 metric.  Same launch shape
// and outputs as k_assign_rgb: grid (64, E), block 256, 4 pixels per thread.
__global__ void __launch_bounds__(256) k_assign_lab(const ImgDev *imgs, const CandEntry *cents, int ncand, int e0, int S,
                                                    int CS, int has_ovr, uint8_t *maps, int to_image, int gi_fmt,
                                                    const TileMove *moves /* per evaluation, or null */) {
    __shared__ float4 pal[MAX_ENTRIES];
    const int e = blockIdx.y, ea = e0 + e, img = ea / ncand, tid = threadIdx.x;
    const ImgDev im = imgs[img];
    const int ovr = has_ovr >= 0 ? cents[ea].slot : -1;   // the entry this evaluation replaces (per evaluation: CandEntry::slot)
    for (int j = tid; j < CS; j += 256) {
        const float *l = (j == ovr) ? cents[ea].lab : im.tables->lab[j];
        pal[j] = make_float4(l[0], l[1], l[2], 0.0f);
    }
    __syncthreads();
    const int q = blockIdx.x * 256 + tid;
    const int px0 = q * 4, y = px0 >> 8, x = px0 & 255;
    const int tile = (y >> 3) * 32 + (x >> 3);
    const int sub = ((moves && moves[ea].tile == tile) ? moves[ea].sub : im.tile_pal[tile]) * S;
    uint32_t packed = 0;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(im.lab) + px0 + k);
        const int a = im.rgba[px0 + k].w;
        // `error < best_error` from best_error = f64::MAX (lib.rs:763, 788): +inf start is equivalent
        // for f32 distances (finite d wins, inf/NaN never does)
        float best = __int_as_float(0x7f800000);
        int bi = 0;
        for (int j = 0; j < S; j++) {
            const float4 c = pal[sub + j];
            const float d = ciede2000(c.x, c.y, c.z, t.x, t.y, t.z);
            if (d < best) {
                best = d;
                bi = j;
            }
        }
        packed |= (uint32_t)(gi_fmt ? (a > 0 ? sub + bi : GI_BLACK) : (a > 0 ? bi : 0)) << (8 * k);
    }
    uint8_t *out = to_image ? im.map : maps + (size_t)e * NPIX;
    reinterpret_cast<uint32_t *>(out)[q] = packed;
}

}  // namespace snes
