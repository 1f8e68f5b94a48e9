// The per-pixel step of k_assign_dither (dither.cuh) in a form that compiles both as device code and as plain host C++.
//
// The device kernel is the product; the host build exists only so that tests/test_dither_core.py can run the very same
// wavefront step -- window rotation, quad fetch, packed keys, exact rounding -- on the CPU, 128 emulated threads at a time,
// and compare it with the oracle where no GPU is present.  Nothing in the library calls the host build.
//
// Reference: optimize() with error diffusion, lib.rs:425-501; color_distance_red_mean lib.rs:1080-1088; lib.rs:773-778 for
// the clamp(0,255).round() of a target.
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__CUDACC__)
#define DC_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define DC_HD inline
#endif

namespace snes {
namespace dc {

constexpr int THREADS = 128;   // thread i owns image rows i and i + 128
constexpr int IW = 256, IH = 256;
constexpr int STEPS = 768;     // wavefront steps t = x + 2y: 766, rounded up to a multiple of the unroll (the last two are idle)
constexpr int GI_TRANSPARENT = 255;

struct alignas(16) U4 {
    uint32_t x, y, z, w;
};
struct alignas(8) U2 {
    uint32_t x, y;
};

// ---- red-mean key in packed form -------------------------------------------------------------------------------------
// key(R,G,B ; r,g,b) = (1024 + r + R)(R - r)^2 + 2048 (G - g)^2 + (1534 - r - R)(B - b)^2      (common.cuh: redmean_key)
//   = C0 + A r - R (r^2 + b^2) - 4096 G g + C1 b + 2B (r b) + T(r,g,b)
//   C0 = (1024 + R) R^2 + 2048 G^2 + (1534 - R) B^2,  A = -R^2 - 2048 R - B^2,  C1 = -2B (1534 - R)
//   T  = r^3 + 1024 r^2 + 2048 g^2 + 1534 b^2 - r b^2        (the same for every entry of a pixel)
// The strict-< first minimum over the entries of a pixel is unchanged by any term that depends on the pixel alone, so the
// kernel evaluates  key' = key - T + L  with  L = KEY_LS (r^2 + b^2) + KEY_LG g + KEY_K0,  whose coefficients fold into the
// entry's table: five multiply-adds per entry.  L is chosen so that key' lies in [-2^28, 2^28) for EVERY pixel colour and
// entry colour (tests/test_dither_core.py checks all 2^24 pixel colours against the per-pixel key range), which leaves room
// for the entry's number in the low three bits:  v = 8 key' + (j & 7)  is an int32, and the signed minimum of v over eight
// consecutive entries is the smallest key AND, among equal keys, the lowest entry -- one VIMNMX per entry instead of a
// compare and two selects.  Two groups of eight cover a 15-colour subpalette.
constexpr int64_t KEY_LS = 1280, KEY_LG = 522240, KEY_K0 = -141273825;

struct alignas(32) KeyCoef {
    U4 a;   // 8 (C0 + K0) + (j & 7),  8 A,  8 (LS - R),  8 (LG - 4096 G)
    U2 b;   // 8 C1,  16 B
    U2 pad;
};
struct alignas(32) PalD {   // as_rgba of an entry as doubles
    double v[4];
};

DC_HD KeyCoef key_coef(int R, int G, int B, int j) {
    // products up to 2^33: 64-bit on the host side of the table, truncated to the 32 bits the device arithmetic keeps
    const int64_t c0 = (int64_t)(1024 + R) * R * R + (int64_t)2048 * G * G + (int64_t)(1534 - R) * B * B + KEY_K0;
    const int64_t a = -(int64_t)R * R - 2048 * (int64_t)R - (int64_t)B * B;
    const int64_t c1 = -2 * (int64_t)B * (1534 - R);
    KeyCoef k;
    k.a.x = (uint32_t)(8 * c0 + (j & 7));
    k.a.y = (uint32_t)(8 * a);
    k.a.z = (uint32_t)(8 * (KEY_LS - R));
    k.a.w = (uint32_t)(8 * (KEY_LG - 4096 * (int64_t)G));
    k.b.x = (uint32_t)(8 * c1);
    k.b.y = (uint32_t)(16 * B);
    k.pad.x = k.pad.y = 0;
    return k;
}

// v = 8 key' + (j & 7) of one entry (mod 2^32 arithmetic; the value itself fits an int32)
DC_HD int packed_key(const KeyCoef &k, uint32_t r, uint32_t g, uint32_t b, uint32_t s2, uint32_t rb) {
    const U4 a = k.a;
    const U2 c = k.b;
    return (int)(a.x + a.y * r + a.z * s2 + a.w * g + c.x * b + c.y * rb);
}

DC_HD int imin(int a, int b) { return a < b ? a : b; }

// index of the first minimum of the red-mean distance among entries tab[0 .. S-1]
DC_HD int nearest_rgb(const KeyCoef *tab, int S, int r, int g, int b) {
    const uint32_t ur = (uint32_t)r, ug = (uint32_t)g, ub = (uint32_t)b, s2 = ur * ur + ub * ub, rb = ur * ub;
#define DC_V(j) packed_key(tab[j], ur, ug, ub, s2, rb)
    if (S == 15) {   // the SNES subpalette: straight-line code
        int m0 = DC_V(0), m1 = DC_V(8);
        m0 = imin(m0, DC_V(1));
        m1 = imin(m1, DC_V(9));
        m0 = imin(m0, DC_V(2));
        m1 = imin(m1, DC_V(10));
        m0 = imin(m0, DC_V(3));
        m1 = imin(m1, DC_V(11));
        m0 = imin(m0, DC_V(4));
        m1 = imin(m1, DC_V(12));
        m0 = imin(m0, DC_V(5));
        m1 = imin(m1, DC_V(13));
        m0 = imin(m0, DC_V(6));
        m1 = imin(m1, DC_V(14));
        m0 = imin(m0, DC_V(7));
        const bool hi = (m1 >> 3) < (m0 >> 3);   // equal keys: the lower group wins
        return hi ? 8 + (m1 & 7) : (m0 & 7);
    }
    if (S == 3) return imin(imin(DC_V(0), DC_V(1)), DC_V(2)) & 7;   // the NES configuration (4 x 3 colours)
    int bestk = 0x7fffffff, bi = 0;
    for (int j0 = 0; j0 < S; j0 += 8) {
        int m = DC_V(j0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int jj = 1; jj < 8; jj++)
            if (j0 + jj < S) m = imin(m, DC_V(j0 + jj));
        const int k = m >> 3;
        if (k < bestk) {
            bestk = k;
            bi = j0 + (m & 7);
        }
    }
    return bi;
#undef DC_V
}

// ---- IEEE helpers: the device intrinsics, or their host equivalents (host translation units are built with
// -ffp-contract=off, so a * b + c below is two roundings there as well) ---------------------------------------------------
DC_HD double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
DC_HD double add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
// a + b rounded toward minus infinity
DC_HD double add_rd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rd(a, b);
#else
    const double s = a + b, t = s - a, e = (a - (s - t)) + (b - t);   // TwoSum: a + b = s + e exactly
    return e < 0.0 ? nextafter(s, -INFINITY) : s;
#endif
}
DC_HD int lo32(double v) {
#if defined(__CUDA_ARCH__)
    return __double2loint(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return (int)(uint32_t)u;
#endif
}
DC_HD double u8_to_double(uint32_t o) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__hiloint2double(0x43300000, (int)o), -4503599627370496.0);   // (2^52 + o) - 2^52, exact: no XU conversion
#else
    return (double)o;
#endif
}
// lib.rs:773-778: target.clamp(0.0, 255.0).round() as u8 -- round() is half away from zero.  For target >= 0 that is
// floor(target + 0.5) with the sum taken EXACTLY; a sum rounded toward minus infinity has the same floor (it never moves up
// past an integer, and an exact sum >= n stays >= n because n is representable).  A negative target ends at 0 under both
// rules once clamped.  floor() of the (small) sum is the low word of its round-down sum with 1.5 * 2^52.
DC_HD int round_clamp_u8(double target) {
    const int v = lo32(add_rd(add_rd(target, 0.5), 6755399441055744.0));
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// ---- per-thread state: registers on the device (every index into win[] is a compile-time constant there) ---------------
struct Thread {
    double win[3][3];   // damped errors of three consecutive pixels of the row above; the roles rotate with the step number
    double ee[3];       // damped error of this row's previous pixel
    uint32_t q[4];      // four source pixels, the current one among them
    uint32_t packed;    // up to four output bytes
};

DC_HD void thread_init(Thread &th) {
    for (int k = 0; k < 3; k++) {
        th.ee[k] = 0.0;
        for (int c = 0; c < 3; c++) th.win[k][c] = 0.0;
    }
    th.packed = 0;
}

// offset of pixel tau (tau = x on row i for tau < 256, x + 256 on row i + 128 afterwards) from the first pixel of row i
DC_HD int pixel_off(int tau) { return tau + (tau >> 8) * (THREADS * IW - IW); }

// where tile (ty, tx)'s subpalette is kept for the thread that owns its rows: [(ty & 15)][ty >> 4][tx], so that thread i at
// wavefront time tau finds it at (i >> 3) * 64 + (tau >> 3)
DC_HD int stp_slot(int tile) {
    const int ty = tile >> 5, tx = tile & 31;
    return (ty & 15) * 64 + (ty >> 4) * 32 + tx;
}

// One wavefront step of thread i at time tau = t - 2 i, ROT = t % 3.
//   mrd: mailbox the thread above wrote in the previous step ([channel * THREADS]); mwr: this thread's slot of this step's mailbox
//   stp: per tile slot (stp_slot), first entry of the tile's subpalette (global entry number)
//   nearest(first_entry, r, g, b) -> index within the subpalette;  pald: as_rgba of every entry as doubles
//   load_quad(pixel_off) -> the four source pixels starting there (RGBA8 words);  out: the row-major output map
// win[(ROT+1)%3], win[(ROT+2)%3], win[ROT] hold the errors of pixels x-1, x, x+1 of the row above once the new value is in.
//   gi_mask: 0xff when the output is the global entry number (255 = transparent), 0 when it is the index within the subpalette
template <int ROT, class Nearest, class LoadQuad>
DC_HD void step(Thread &th, int tau, int i, const double *mrd, double *mwr, const uint8_t *stp, const PalD *pald, uint32_t gi_mask,
                uint8_t *out_row_i, Nearest nearest, LoadQuad load_quad) {
    constexpr int P = ROT % 3, Q = (ROT + 1) % 3, R = (ROT + 2) % 3;
    // the row above finished its pixel x+1 in the previous step (for x = 255 that is already pixel 0 of the row after it; the
    // term is weighted out below).  Reading every step, busy or not, keeps the rotation unconditional; what an idle thread
    // reads is finite and never used with a non-zero weight.  Image row 0 has no row above and needs no test either: thread
    // 0's "row above" is thread 127, whose first write happens at step 254 (its pixel 0) and is read by thread 0 at step 255,
    // i.e. as pixel x+1 = 0 of row 127 for pixel 255 of row 0 -- weighted out -- and then for row 128; before that the slot
    // still holds the zeros it was initialised with.
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 3; c++) th.win[P][c] = mrd[c * THREADS];
    if ((unsigned)tau >= 2u * IW) return;
    const int x = tau & (IW - 1);
    const uint32_t lo = (tau & 1) ? th.q[1] : th.q[0], hi = (tau & 1) ? th.q[3] : th.q[2];
    const uint32_t pw = (tau & 2) ? hi : lo;
    if ((tau & 3) == 3 && tau + 1 < 2 * IW) load_quad(pixel_off(tau + 1), th.q);   // lands during this step, used from the next
    const int first = stp[(i >> 3) * 64 + (tau >> 3)];
    // A term the reference skips (lib.rs:478-493: x + 1 < width, x > 0) gets weight zero: it contributes +-0, which leaves the
    // running sum unchanged, and the sum's leading `0.0 +` only ever changes the sign of a zero, which nothing downstream sees.
    const double wse = x > 0 ? 1.0 / 16.0 : 0.0, wsw = x + 1 < IW ? 3.0 / 16.0 : 0.0, we = x > 0 ? 7.0 / 16.0 : 0.0;
    double acc[3], target[3];
    int t8[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 3; c++) {
        // every stored error is already damped (e * 0.8, the first product of each term of lib.rs:479-493); the terms are added
        // in the raster order of the contributing pixels: SE, S, SW, then E
        double a = mul(th.win[Q][c], wse);
        a = add(a, mul(th.win[R][c], 5.0 / 16.0));
        a = add(a, mul(th.win[P][c], wsw));
        a = add(a, mul(th.ee[c], we));
        acc[c] = a;
        target[c] = add(u8_to_double((pw >> (8 * c)) & 255u), a);
        t8[c] = round_clamp_u8(target[c]);
    }
    uint32_t idx;
    if (pw >> 24) {
        const int bi = nearest(first, t8[0], t8[1], t8[2]);
        const PalD nc = pald[first + bi];
        th.ee[0] = mul(add(target[0], -nc.v[0]), 0.8);
        th.ee[1] = mul(add(target[1], -nc.v[1]), 0.8);
        th.ee[2] = mul(add(target[2], -nc.v[2]), 0.8);
        idx = ((uint32_t)first & gi_mask) + (uint32_t)bi;
    } else {   // transparent: the accumulated error moves on unchanged, index 0 (lib.rs:453-475)
        th.ee[0] = mul(acc[0], 0.8);
        th.ee[1] = mul(acc[1], 0.8);
        th.ee[2] = mul(acc[2], 0.8);
        idx = gi_mask;   // GI_TRANSPARENT or 0
    }
    mwr[0] = th.ee[0];
    mwr[THREADS] = th.ee[1];
    mwr[2 * THREADS] = th.ee[2];
    th.packed |= idx << (8 * (tau & 3));
    if ((tau & 3) == 3) {
        *reinterpret_cast<uint32_t *>(out_row_i + pixel_off(tau - 3)) = th.packed;
        th.packed = 0;
    }
}

}  // namespace dc
}  // namespace snes
