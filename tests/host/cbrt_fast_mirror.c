/* Host mirror of msun_cbrtf_fast (snesimage_b200/csrc/common.cuh) -- test infrastructure, built by tests/test_cbrt_fast.py.
 * The device function evaluates yuvxyb-math's cbrtf (FreeBSD msun) as one f32 and one division-free f64 Halley step and calls the
 * exact function where the f64 result lies within 2^-39 of an f32 rounding boundary.  This file restates that arithmetic in plain C
 * (same operations, same integer widening / narrowing; the two hardware approximations -- __fdividef and rcp.approx -- are the
 * correctly rounded operations here, perturbed by `pert` ulps to cover their error bounds) so that the ALGORITHM -- window, integer
 * rounding, fallback rule -- can be checked against the oracle's cbrtf on the CPU.  The device code itself is checked on the GPU
 * (snes_ctx_cbrt_selfcheck).  Compile with -ffp-contract=off. */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef float (*exact_fn)(float);

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline double widen_pos(float f) {
    const uint32_t b = f2u(f);
    const uint64_t u = ((uint64_t)((b >> 3) + 0x38000000u) << 32) | (uint64_t)(uint32_t)(b << 29);
    double d; memcpy(&d, &u, 8); return d;
}
static inline uint32_t narrow_pos_bits(double d) {
    uint64_t u; memcpy(&u, &d, 8);
    return ((((uint32_t)(u >> 32)) - 0x38000000u) << 3) | ((uint32_t)u >> 29);
}

static float fast(float x, int pert, exact_fn exact, long *fallbacks) {
    const uint32_t bits = f2u(x);
    if (bits - 0x00800000u >= 0x7b000000u) return exact(x);
    const float t0 = u2f(bits / 3 + 709958130u);
    const float r0 = (t0 * t0) * t0;
    float q = ((x + x) + r0) / ((x + r0) + r0);
    q = u2f(f2u(q) + (uint32_t)pert);
    const float t1 = t0 * q;
    const double t = widen_pos(t1), xd = widen_pos(x);
    const double r = t * t * t;
    const double num = xd + xd + r, den = xd + r + r;
    float rf = 1.0f / u2f(narrow_pos_bits(den));
    rf = u2f(f2u(rf) - (uint32_t)pert);
    double rd = widen_pos(rf);
    rd = fma(rd, fma(-den, rd, 1.0), rd);
    const double t2 = (t * num) * rd;
    uint64_t u; memcpy(&u, &t2, 8);
    const uint32_t lo = (uint32_t)u;
    int d = (int)(lo & 0x1fffffffu) - (1 << 28);
    if (d < 0) d = -d;
    if (d < (1 << 14)) { (*fallbacks)++; return exact(x); }
    return u2f(narrow_pos_bits(t2) + ((lo >> 28) & 1u));
}

/* every `stride`-th float with bit pattern in [lo_bits, hi_bits) */
void cbrt_compare(uint32_t lo_bits, uint32_t hi_bits, uint32_t stride, int pert, exact_fn exact, long *mismatches, long *fallbacks, long *count) {
    *mismatches = *fallbacks = *count = 0;
    for (uint64_t b = lo_bits; b < hi_bits; b += stride) {
        const float x = u2f((uint32_t)b);
        (*count)++;
        if (f2u(fast(x, pert, exact, fallbacks)) != f2u(exact(x))) (*mismatches)++;
    }
}
