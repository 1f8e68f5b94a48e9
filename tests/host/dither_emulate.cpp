// Host emulation of k_assign_dither's wavefront (test infrastructure; built by tests/test_dither_core.py with g++ -ffp-contract=off).
// It runs snesimage_b200/csrc/dither_core.h's dc::step -- the code the device kernel runs -- for 128 emulated threads, step by
// step, with the same mailbox ring, so that the window rotation, quad fetch, packed keys and exact rounding can be compared
// with the oracle's optimize() (lib.rs:425-501) where no GPU is present.  The library never links or calls this file.
#include <vector>

#include "../../snesimage_b200/csrc/dither_core.h"

using namespace snes::dc;

extern "C" {

// rgba: 256*256*4, tile_sub: 1024 (tile_palettes), rgb8: C*S*3 as_rgba of every entry, out: 256*256
void dither_emulate(const uint8_t *rgba, const uint8_t *tile_sub, const uint8_t *rgb8, int C, int S, int gi_fmt, uint8_t *out) {
    const int CS = C * S;
    std::vector<KeyCoef> ktab(CS);
    std::vector<PalD> pald(CS);
    for (int j = 0; j < CS; j++) {
        ktab[j] = key_coef(rgb8[3 * j], rgb8[3 * j + 1], rgb8[3 * j + 2], j % S);
        pald[j].v[0] = rgb8[3 * j];
        pald[j].v[1] = rgb8[3 * j + 1];
        pald[j].v[2] = rgb8[3 * j + 2];
        pald[j].v[3] = 0.0;
    }
    std::vector<uint8_t> stp(1024);
    for (int j = 0; j < 1024; j++) stp[stp_slot(j)] = (uint8_t)(tile_sub[j] * S);
    static double mail[3][3][THREADS];
    memset(mail, 0, sizeof(mail));
    std::vector<Thread> th(THREADS);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(rgba);
    for (int i = 0; i < THREADS; i++) {
        thread_init(th[i]);
        for (int k = 0; k < 4; k++) th[i].q[k] = src[i * IW + k];
    }
    auto nearest = [&](int first, int r, int g, int b) { return nearest_rgb(ktab.data() + first, S, r, g, b); };
    for (int t = 0; t < STEPS; t++) {
        for (int i = 0; i < THREADS; i++) {
            const int up = (i + THREADS - 1) & (THREADS - 1), tau = t - 2 * i;
            auto load_quad = [&](int off, uint32_t (&q)[4]) {
                for (int k = 0; k < 4; k++) q[k] = src[i * IW + off + k];
            };
            const double *mrd = &mail[(t + 2) % 3][0][up];
            double *mwr = &mail[t % 3][0][i];
            uint8_t *row = out + i * IW;
            switch (t % 3) {
                case 0: step<0>(th[i], tau, i, mrd, mwr, stp.data(), pald.data(), gi_fmt ? 0xffu : 0u, row, nearest, load_quad); break;
                case 1: step<1>(th[i], tau, i, mrd, mwr, stp.data(), pald.data(), gi_fmt ? 0xffu : 0u, row, nearest, load_quad); break;
                default: step<2>(th[i], tau, i, mrd, mwr, stp.data(), pald.data(), gi_fmt ? 0xffu : 0u, row, nearest, load_quad); break;
            }
        }
    }
}

// nearest_rgb on its own: n targets against one subpalette of S colours
void nearest_many(const uint8_t *pal_rgb8, int S, const uint8_t *targets, int n, int32_t *out) {
    std::vector<KeyCoef> ktab(S);
    for (int j = 0; j < S; j++) ktab[j] = key_coef(pal_rgb8[3 * j], pal_rgb8[3 * j + 1], pal_rgb8[3 * j + 2], j);
    for (int k = 0; k < n; k++) out[k] = nearest_rgb(ktab.data(), S, targets[3 * k], targets[3 * k + 1], targets[3 * k + 2]);
}

void round_many(const double *v, int n, int32_t *out) {
    for (int k = 0; k < n; k++) out[k] = round_clamp_u8(v[k]);
}

void key_constants(int64_t *out) {
    out[0] = KEY_LS;
    out[1] = KEY_LG;
    out[2] = KEY_K0;
}
}
