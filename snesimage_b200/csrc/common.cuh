// Shared device-side definitions for libsnesgpu (sm_100a).
//
// Numerics contract: this translation unit is compiled with -fmad=false, so every f32/f64 operation
// below is the IEEE operation as written; fused multiply-adds appear only where __fmaf_rn()/fma() is
// spelled out (where ssimulacra2/yuvxyb use mul_add).  With the same operation order as the CPU
// oracle, all f32 planes are bit-identical to it; only the order of the final f64 sums differs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace snes {

constexpr int W = 256, H = 256;
constexpr int NPIX = W * H;
constexpr int NTILES = 1024;
constexpr int NSCALES = 6;
constexpr int TOTPIX = 87360;              // 65536+16384+4096+1024+256+64
constexpr int EVAL_XYB_FLOATS = 3 * TOTPIX;  // positive-XYB pyramid of one evaluation
constexpr int NSUMS = 6;                   // ssim d, d^4; edge artifact, artifact^4, detail_lost, detail_lost^4
constexpr int PART_DOUBLES = NSCALES * 3 * NSUMS;  // partial sums of one evaluation: [scale][channel][NSUMS]
constexpr int MAX_ENTRIES = 256;           // sub_count * sub_size
constexpr int BLACK = 256;                 // table slot of a transparent (rendered black) pixel
constexpr int GI_BLACK = 255;              // transparent pixel in a gi-format scratch map (needs C*S <= 255)
constexpr int NES_COUNT = 56;

__host__ __device__ __forceinline__ int scale_off(int s) {  // pixel offset of scale s inside a pyramid
    // sum_{t<s} (256>>t)^2 = 65536 * (1 - 4^-s) * 4/3 : 0, 65536, 81920, 86016, 87040, 87296, 87360
    return (262144 - (262144 >> (2 * s))) / 3;
}

// Per-image base tables for every palette entry (+ slot 256 = black), rebuilt whenever the palette changes.
struct PalTables {
    float lin[MAX_ENTRIES + 1][3];  // linear RGB of as_rgba(entry)
    float xyb[MAX_ENTRIES + 1][3];  // positive XYB of the same
    uchar4 rgb8[MAX_ENTRIES];       // SnesColor::as_rgba (lib.rs:662-669), u8-wrapping
    float lab[MAX_ENTRIES][4];      // Lab<D65,f32> of rgb8 (perceptual mode)
};

// The one palette entry a candidate replaces: its colour in every form the kernels need, and WHICH entry it replaces
// (slot = palette * S + index).  The slot travels with the evaluation, so one launch can carry candidates of several
// entries (a whole sweep, or the speculated next iterations of one picture).
struct CandEntry {
    float lin[3];
    float xyb[3];
    uchar4 rgb8;
    float lab[3];
    int32_t slot;
    uint32_t pad;
};

// Device view of one OptimizedImage.
struct ImgDev {
    const uchar4 *rgba;       // original, row-major
    uint8_t *tile_pal;        // 1024
    uint8_t *palette;         // C*S*3, "5-bit" values
    uint8_t *map;             // palette_map
    const float *xyb_rm;      // source positive-XYB pyramid [scale][ch][y][x]
    const float *xyb_cm;      // same, transposed [scale][ch][x][y]
    float *mu1;               // blur(i1)       [scale][ch][y][x]
    float *s11;               // blur(i1*i1)    [scale][ch][y][x]
    float2 *ms11;             // (mu1, s11) interleaved, same index space (read by the scorers' maps)
    float2 *bfxb;             // scale 0 only, [y][x]: (|i1 - mu1| of channel X, of channel B): the source side of edge_diff_map
                              // for the two channels whose ssim_map carries no weight at scale 0 (score_v3.cuh: v3_scale0_pair)
    PalTables *tables;
    double *cur_err;          // error() of the current state
    const float *lab;         // per-pixel Lab of the original (perceptual mode), [NPIX][4]
    const uint8_t *alpha;     // alpha channel of the original, [NPIX]
    // per-step scratch of the no-dither candidate path (assign_delta.cuh), independent of the entry a candidate replaces
    uint8_t *base_gi;         // [NPIX] assignment under the current palette as global entry index (GI_BLACK: transparent)
    uint8_t *sec_idx;         // [NPIX] runner-up: first minimum over the entries other than the pixel's best (0xFF: S == 1)
    int2 *keys;               // [NPIX] (key of the best, key of the runner-up): int32 red-mean keys, or f32 CIEDE2000 bits
};

// A tile-reassignment candidate: tile `tile` (tile_y * 32 + tile_x) bound to subpalette `sub` instead of
// tile_palettes[tile] (what a click on the tile does, lib.rs:1005-1017, as an evaluated candidate; TODO.md:36-37).
struct TileMove {
    int32_t tile, sub;
};

struct Best {
    double err;
    int32_t idx;
    int32_t pad;
};

// ---- colour helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uchar4 snes_as_rgba(uint8_t r, uint8_t g, uint8_t b) {  // lib.rs:662-669
    return make_uchar4((uint8_t)((uint8_t)(r * 8) + r / 4), (uint8_t)((uint8_t)(g * 8) + g / 4),
                       (uint8_t)((uint8_t)(b * 8) + b / 4), 255);
}

// lib.rs:1080-1088 as an order-equivalent int32 key (= 512 * distance^2): the f64 expression is exact
// for u8 operands and sqrt is monotone, so `<` and ties on the key are those of the reference.
__device__ __forceinline__ int redmean_key(int r1, int g1, int b1, int r2, int g2, int b2) {
    const int rs = r1 + r2, dr = r1 - r2, dg = g1 - g2, db = b1 - b2;
    return (1024 + rs) * dr * dr + 2048 * dg * dg + (1534 - rs) * db * db;
}

// FreeBSD msun s_cbrtf.c (as ported by yuvxyb-math): bit-exact restatement, f64 Halley steps.
__device__ __forceinline__ float msun_cbrtf(float x) {
    const uint32_t B1 = 709958130u, B2 = 642849266u;
    uint32_t bits = __float_as_uint(x);
    uint32_t hx = bits & 0x7fffffffu;
    const uint32_t sign = bits & 0x80000000u;
    if (hx >= 0x7f800000u) return x + x;
    if (hx < 0x00800000u) {
        if (hx == 0) return x;
        hx = __float_as_uint(x * 16777216.0f) & 0x7fffffffu;
        hx = hx / 3 + B2;
    } else {
        hx = hx / 3 + B1;
    }
    double t = (double)__uint_as_float(sign | hx);
    const double xd = (double)x;
    double r = t * t * t;
    t = t * (xd + xd + r) / (xd + r + r);
    r = t * t * t;
    t = t * (xd + xd + r) / (xd + r + r);
    return (float)t;
}

// The same function, ~2.5x cheaper, bit for bit: msun's result is (float)t2 with t2 = two f64 Halley steps from a 5-bit guess, i.e.
// the f64 value cbrt(x)(1 + e), |e| < 2^-46.  Here the first Halley step runs in f32 with an approximate division (15 good bits
// are all it can give either way) and the second in f64 with the division replaced by a Newton-refined reciprocal; the result
// t2' differs from t2 by less than 2^-43 relative (measured over every float in [2^-9, 2) with both approximations perturbed by
// +-2 ulp: 2^-43.1; tests/test_gpu_parity.py::test_fast_cbrt_is_the_exact_one compares the two functions on the GPU over the same
// range).  (float)t2' can differ from (float)t2 only if a rounding boundary of f32 -- a midpoint between two floats -- lies between
// them, so whenever t2' is within 2^-39 of a midpoint (bits 28..0 of its mantissa within 2^14 of 0x10000000: 6 inputs in 100,000)
// the exact function is evaluated instead.  What it saves is the FP64 pipe: two f64 divisions and ~60 FP64 instructions per cube
// root become ~10 (k_assign_pyr<2> converts 84 M pixels per 4096 dithered evaluations), and its five f32 <-> f64 conversions leave
// the XU pipe (integer widening / narrowing of positive normal numbers), which the reciprocals need.
#ifndef SNES_EXACT_CBRT
#define SNES_EXACT_CBRT 0   // 1: every cube root through msun_cbrtf (A/B runs)
#endif
// (the rare path as a real call: three inlined copies of the exact function per pixel conversion cost more in code size and
// registers than the fast path saves in CIELAB mode)
__device__ __noinline__ float msun_cbrtf_call(float x) { return msun_cbrtf(x); }
// f32 <-> f64 of positive normal numbers by integer operations (exact widening; narrowing toward zero): the conversion
// instructions share the quarter-rate XU pipe with the two reciprocals below, which is what bounds the conversion kernels
__device__ __forceinline__ double widen_pos(float f) {
    const uint32_t b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
__device__ __forceinline__ uint32_t narrow_pos_bits(double d) {
    return (((uint32_t)__double2hiint(d) - 0x38000000u) << 3) | ((uint32_t)__double2loint(d) >> 29);
}
__device__ __forceinline__ float msun_cbrtf_fast(float x, unsigned *fallbacks = nullptr) {
#if SNES_EXACT_CBRT
    return msun_cbrtf(x);
#else
    const uint32_t bits = __float_as_uint(x);
    if (bits - 0x00800000u >= 0x7b000000u) return msun_cbrtf_call(x);   // zero, subnormal, >= 2^120 (the reciprocal of 3x must stay a normal float), infinite, NaN, negative
    const float t0 = __uint_as_float(bits / 3 + 709958130u);
    const float r0 = (t0 * t0) * t0;
    const float t1 = t0 * __fdividef((x + x) + r0, (x + r0) + r0);
    const double t = widen_pos(t1), xd = widen_pos(x);
    const double r = t * t * t;
    const double num = xd + xd + r, den = xd + r + r;
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(__uint_as_float(narrow_pos_bits(den))));
    double rd = widen_pos(rf);
    rd = fma(rd, fma(-den, rd, 1.0), rd);
    const double t2 = (t * num) * rd;
    const uint32_t lo = (uint32_t)__double2loint(t2);
    if (abs((int)(lo & 0x1fffffffu) - (1 << 28)) < (1 << 14)) {
        if (fallbacks) atomicAdd(fallbacks, 1u);
        return msun_cbrtf_call(x);
    }
    // (float)t2, round to nearest: t2 is not within 2^14 units of a tie, so adding the first dropped bit is the whole rule
    return __uint_as_float(narrow_pos_bits(t2) + ((lo >> 28) & 1u));
#endif
}

// yuvxyb linear_rgb_to_xyb + ssimulacra2 make_positive_xyb
__device__ __forceinline__ void lin_to_pxyb(float r, float g, float b, float &X, float &Y, float &B) {
    const float kM02 = 0.078f, kM00 = 0.30f, kM01 = 1.0f - kM02 - kM00;
    const float kM12 = 0.078f, kM10 = 0.23f, kM11 = 1.0f - kM12 - kM10;
    const float kM20 = 0.24342269f, kM21 = 0.20476745f, kM22 = 1.0f - kM20 - kM21;
    const float kB0 = 0.0037930734f;
    const float kNegBias = -0.15595420f;
    float m0 = __fmaf_rn(kM00, r, __fmaf_rn(kM01, g, __fmaf_rn(kM02, b, kB0)));
    float m1 = __fmaf_rn(kM10, r, __fmaf_rn(kM11, g, __fmaf_rn(kM12, b, kB0)));
    float m2 = __fmaf_rn(kM20, r, __fmaf_rn(kM21, g, __fmaf_rn(kM22, b, kB0)));
    m0 = m0 < 0.0f ? 0.0f : m0;
    m1 = m1 < 0.0f ? 0.0f : m1;
    m2 = m2 < 0.0f ? 0.0f : m2;
    m0 = msun_cbrtf_fast(m0) + kNegBias;
    m1 = msun_cbrtf_fast(m1) + kNegBias;
    m2 = msun_cbrtf_fast(m2) + kNegBias;
    const float x = 0.5f * (m0 - m1), y = 0.5f * (m0 + m1);
    B = (m2 - y) + 0.55f;
    X = __fmaf_rn(x, 14.0f, 0.42f);
    Y = y + 0.01f;
}

}  // namespace snes
