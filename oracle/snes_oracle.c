/*
 * oracle/snes_oracle.c  --  TEST INFRASTRUCTURE: CPU restatement of the snesimage hot path.
 *
 * NOT product code.  Single-threaded plain C, built with  gcc -O2 -ffp-contract=off  (no
 * -ffast-math) so that every f32/f64 operation is the IEEE operation written here; fused
 * multiply-adds appear only where fmaf()/fma() is spelled out (where the crates use mul_add).
 *
 * First-party arithmetic follows /root/reference/src/lib.rs line by line (cited per function) and is
 * pinned by that file alone.  Third-party arithmetic (cogset, palette, ssimulacra2/yuvxyb) is a
 * restatement of the published algorithms -- PARITY UNPINNED against the crates themselves, see
 * constants_unverified.h.
 */
#include "snes_oracle.h"
#include "constants_unverified.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ============================================================================================ */
/* state: struct OptimizedImage (lib.rs:33-43), struct Palette (lib.rs:747-751)                   */
/* ============================================================================================ */
struct ora_image {
    int width, height;
    uint8_t *original;      /* N x RGBA8 */
    uint8_t tile_palettes[ORA_NTILES];
    uint8_t *colors;        /* sub_count*sub_size x 3, "5-bit" values (may hold 32, see lib.rs:396) */
    int sub_count, sub_size;
    uint8_t *palette_map;   /* N */
    int dither, perceptual_palettes, nes;
};

ora_image *ora_image_new(const uint8_t *rgba, int width, int height, int sub_count, int sub_size, int dither,
                         int perceptual_palettes, int nes) {
    /* The reference's size check (lib.rs:838-840) uses && and lets 256xN through, after which its
     * fixed 32x32 tile table (lib.rs:58, 565) is wrong; the restatement covers 256x256 only. */
    if (width != ORA_WIDTH || height != ORA_HEIGHT || sub_count < 1 || sub_size < 1 || sub_count * sub_size > 256)
        return NULL;
    ora_image *im = (ora_image *)calloc(1, sizeof(ora_image));
    im->width = width;
    im->height = height;
    im->original = (uint8_t *)malloc((size_t)ORA_NPIX * 4);
    memcpy(im->original, rgba, (size_t)ORA_NPIX * 4);
    im->colors = (uint8_t *)calloc((size_t)sub_count * sub_size, 3); /* Palette::new: all (0,0,0) */
    im->sub_count = sub_count;
    im->sub_size = sub_size;
    im->palette_map = (uint8_t *)calloc(ORA_NPIX, 1);
    im->dither = dither;
    im->perceptual_palettes = perceptual_palettes;
    im->nes = nes;
    return im;
}

void ora_image_free(ora_image *im) {
    if (!im) return;
    free(im->original);
    free(im->colors);
    free(im->palette_map);
    free(im);
}

void ora_get_palette(const ora_image *im, uint8_t *out) { memcpy(out, im->colors, (size_t)im->sub_count * im->sub_size * 3); }
void ora_set_palette(ora_image *im, const uint8_t *in) { memcpy(im->colors, in, (size_t)im->sub_count * im->sub_size * 3); }
void ora_get_tile_palettes(const ora_image *im, uint8_t *out) { memcpy(out, im->tile_palettes, ORA_NTILES); }
void ora_set_tile_palettes(ora_image *im, const uint8_t *in) { memcpy(im->tile_palettes, in, ORA_NTILES); }
void ora_get_palette_map(const ora_image *im, uint8_t *out) { memcpy(out, im->palette_map, ORA_NPIX); }
void ora_set_palette_map(ora_image *im, const uint8_t *in) { memcpy(im->palette_map, in, ORA_NPIX); }

static inline const uint8_t *orig_px(const ora_image *im, int x, int y) { /* lib.rs:75-77 */
    return im->original + 4 * ((size_t)y * im->width + x);
}

/* ============================================================================================ */
/* SnesColor (lib.rs:628-745)                                                                     */
/* ============================================================================================ */
static const uint8_t NES_TABLE[ORA_NES_COLOR_COUNT][3] = { /* lib.rs:687-742 */
    {13, 13, 13}, {0, 2, 16},   {3, 0, 17},   {7, 0, 15},   {10, 0, 10},  {11, 0, 3},   {9, 2, 0},    {7, 3, 0},
    {4, 6, 0},    {0, 7, 0},    {0, 8, 0},    {0, 7, 4},    {0, 5, 10},   {0, 0, 0},    {23, 23, 23}, {3, 10, 24},
    {9, 6, 28},   {14, 4, 26},  {18, 3, 21},  {19, 5, 11},  {19, 6, 0},   {15, 9, 0},   {11, 12, 0},  {4, 14, 0},
    {0, 15, 0},   {0, 14, 8},   {0, 13, 17},  {0, 0, 0},    {31, 31, 31}, {13, 20, 31}, {17, 19, 31}, {22, 16, 31},
    {27, 14, 31}, {28, 14, 23}, {28, 17, 13}, {26, 19, 5},  {22, 21, 1},  {15, 24, 2},  {10, 25, 8},  {8, 25, 16},
    {8, 24, 24},  {9, 9, 9},    {31, 31, 31}, {25, 29, 31}, {27, 27, 31}, {29, 27, 31}, {31, 26, 31}, {31, 26, 30},
    {31, 27, 25}, {31, 28, 22}, {30, 30, 21}, {27, 31, 21}, {25, 31, 23}, {24, 31, 26}, {24, 30, 30}, {23, 24, 23}};

void ora_nes_color(int index, uint8_t out[3]) {
    if (index >= 0 && index < ORA_NES_COLOR_COUNT) {
        memcpy(out, NES_TABLE[index], 3);
    } else {
        out[0] = out[1] = out[2] = 0; /* lib.rs:743 */
    }
}

/* lib.rs:662-669.  u8 arithmetic; release builds wrap (c == 32 gives 8), debug builds panic. */
void ora_snes_as_rgba(const uint8_t c5[3], uint8_t out[4]) {
    for (int i = 0; i < 3; i++) out[i] = (uint8_t)((uint8_t)(c5[i] * 8) + c5[i] / 4);
    out[3] = 255;
}

uint16_t ora_snes_as_u16(const uint8_t c5[3]) { /* lib.rs:679-681 */
    return (uint16_t)((uint16_t)c5[0] + ((uint16_t)c5[1] << 5) + ((uint16_t)c5[2] << 10));
}

/* ============================================================================================ */
/* palette 0.7.6 restated: Srgb<u8> -> Lab<D65,f32>, Lab<f64> -> Srgb<u8>, CIEDE2000              */
/* ============================================================================================ */
/* The two sRGB -> linear transfer functions only ever see 8-bit inputs on this path, so each is a 256-entry table.  The
 * built-in tables come from the recalled formulas with libm's powf; ora_set_transfer_luts() swaps in tables dumped from the
 * real crates (tests/golden/gen_reference_vectors.rs), which is how a maintainer with cargo pins this oracle. */
static float g_lut_palette[256], g_lut_yuvxyb[256];
static int g_luts_ready = 0;
static float srgb_eotf_formula(float x);
static void luts_init(void) {
    if (g_luts_ready) return;
    for (int v = 0; v < 256; v++) {
        const float c = (float)v / 255.0f; /* into_format */
        g_lut_palette[v] = c <= 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f); /* Srgb::into_linear */
        g_lut_yuvxyb[v] = srgb_eotf_formula(c);
    }
    g_luts_ready = 1;
}
void ora_set_transfer_luts(const float *yuvxyb_eotf, const float *palette_eotf) {
    g_luts_ready = 0;
    luts_init();
    if (yuvxyb_eotf) memcpy(g_lut_yuvxyb, yuvxyb_eotf, sizeof g_lut_yuvxyb);
    if (palette_eotf) memcpy(g_lut_palette, palette_eotf, sizeof g_lut_palette);
}
void ora_get_transfer_luts(float *yuvxyb_eotf, float *palette_eotf) {
    luts_init();
    memcpy(yuvxyb_eotf, g_lut_yuvxyb, sizeof g_lut_yuvxyb);
    memcpy(palette_eotf, g_lut_palette, sizeof g_lut_palette);
}

void ora_srgb8_to_lab_f32(uint8_t r8, uint8_t g8, uint8_t b8, float out[3]) {
    luts_init();
    float c[3] = {g_lut_palette[r8], g_lut_palette[g8], g_lut_palette[b8]}; /* into_format + Srgb::into_linear */
    /* multiply_rgb_to_xyz: (m0*r + m1*g) + m2*b */
    float x = ((float)ORA_XYZ_M00 * c[0] + (float)ORA_XYZ_M01 * c[1]) + (float)ORA_XYZ_M02 * c[2];
    float y = ((float)ORA_XYZ_M10 * c[0] + (float)ORA_XYZ_M11 * c[1]) + (float)ORA_XYZ_M12 * c[2];
    float z = ((float)ORA_XYZ_M20 * c[0] + (float)ORA_XYZ_M21 * c[1]) + (float)ORA_XYZ_M22 * c[2];
    x /= (float)ORA_D65_X;
    y /= (float)ORA_D65_Y;
    z /= (float)ORA_D65_Z;
    const float eps = (float)((6.0 / 29.0) * (6.0 / 29.0) * (6.0 / 29.0));
    const float kappa = (float)(841.0 / 108.0);
    const float delta = (float)(4.0 / 29.0);
    float fx = x > eps ? cbrtf(x) : kappa * x + delta;
    float fy = y > eps ? cbrtf(y) : kappa * y + delta;
    float fz = z > eps ? cbrtf(z) : kappa * z + delta;
    out[0] = fy * 116.0f - 16.0f;
    out[1] = (fx - fy) * 500.0f;
    out[2] = (fy - fz) * 200.0f;
}

static uint8_t f64_to_u8_stimulus(double v) { /* palette IntoStimulus<u8> for f64: clamp, round-half-even */
    double s = v * 255.0;
    if (!(s > 0.0)) return 0; /* also NaN */
    if (s > 255.0) s = 255.0;
    return (uint8_t)nearbyint(s);
}

void ora_lab_f64_to_srgb8(const double lab[3], uint8_t out[3]) {
    double fy = (lab[0] + 16.0) / 116.0;
    double fx = fy + lab[1] / 500.0;
    double fz = fy - lab[2] / 200.0;
    const double eps = 6.0 / 29.0, kappa = 108.0 / 841.0, delta = 4.0 / 29.0;
    double x = (fx > eps ? fx * fx * fx : (fx - delta) * kappa) * ORA_D65_X;
    double y = (fy > eps ? fy * fy * fy : (fy - delta) * kappa) * ORA_D65_Y;
    double z = (fz > eps ? fz * fz * fz : (fz - delta) * kappa) * ORA_D65_Z;
    double lin[3];
    lin[0] = (ORA_RGB_M00 * x + ORA_RGB_M01 * y) + ORA_RGB_M02 * z;
    lin[1] = (ORA_RGB_M10 * x + ORA_RGB_M11 * y) + ORA_RGB_M12 * z;
    lin[2] = (ORA_RGB_M20 * x + ORA_RGB_M21 * y) + ORA_RGB_M22 * z;
    for (int i = 0; i < 3; i++) {
        double v = lin[i] <= 0.0031308 ? 12.92 * lin[i] : 1.055 * pow(lin[i], 1.0 / 2.4) - 0.055;
        out[i] = f64_to_u8_stimulus(v);
    }
}

#define CIEDE2000_BODY(T, SQRT, ATAN2, SIN, COS, EXP, FABS)                                              \
    const T pi_over_180 = (T)(3.14159265358979323846 / 180.0);                                          \
    const T twenty_five_pow_seven = (T)6103515625.0;                                                    \
    T c1 = SQRT(lab1[1] * lab1[1] + lab1[2] * lab1[2]);                                                 \
    T c2 = SQRT(lab2[1] * lab2[1] + lab2[2] * lab2[2]);                                                 \
    T delta_l_prime = lab2[0] - lab1[0];                                                                \
    T l_bar = (lab1[0] + lab2[0]) / (T)2;                                                               \
    T c_bar = (c1 + c2) / (T)2;                                                                         \
    T c_bar2 = c_bar * c_bar;                                                                           \
    T c_bar7 = c_bar2 * c_bar2 * c_bar2 * c_bar;                                                        \
    T g = (T)0.5 * ((T)1 - SQRT(c_bar7 / (c_bar7 + twenty_five_pow_seven)));                            \
    T a1p = lab1[1] * ((T)1 + g);                                                                       \
    T a2p = lab2[1] * ((T)1 + g);                                                                       \
    T c1p = SQRT(a1p * a1p + lab1[2] * lab1[2]);                                                        \
    T c2p = SQRT(a2p * a2p + lab2[2] * lab2[2]);                                                        \
    T h1p = 0, h2p = 0;                                                                                 \
    if (!(lab1[2] == 0 && a1p == 0)) {                                                                  \
        h1p = ATAN2(lab1[2], a1p) / pi_over_180;                                                        \
        if (h1p < 0) h1p += (T)360;                                                                     \
    }                                                                                                   \
    if (!(lab2[2] == 0 && a2p == 0)) {                                                                  \
        h2p = ATAN2(lab2[2], a2p) / pi_over_180;                                                        \
        if (h2p < 0) h2p += (T)360;                                                                     \
    }                                                                                                   \
    T h_diff = h2p - h1p;                                                                               \
    T h_abs = FABS(h_diff);                                                                             \
    int zero_chroma = (c1p == 0 || c2p == 0);                                                           \
    T delta_h_prime;                                                                                    \
    if (zero_chroma) delta_h_prime = 0;                                                                 \
    else if (h_abs <= (T)180) delta_h_prime = h_diff;                                                   \
    else if (h2p <= h1p) delta_h_prime = h_diff + (T)360;                                               \
    else delta_h_prime = h_diff - (T)360;                                                               \
    T delta_big_h = (T)2 * SQRT(c1p * c2p) * SIN(delta_h_prime / (T)2 * pi_over_180);                   \
    T h_bar;                                                                                            \
    if (zero_chroma) h_bar = h1p + h2p;                                                                 \
    else if (h_abs > (T)180) {                                                                          \
        if (h1p + h2p < (T)360) h_bar = (h1p + h2p + (T)360) / (T)2;                                    \
        else h_bar = (h1p + h2p - (T)360) / (T)2;                                                       \
    } else h_bar = (h1p + h2p) / (T)2;                                                                  \
    T lb50 = (l_bar - (T)50) * (l_bar - (T)50);                                                         \
    T c_bar_p = (c1p + c2p) / (T)2;                                                                     \
    T t = (T)1 - (T)0.17 * COS((h_bar - (T)30) * pi_over_180) + (T)0.24 * COS(((T)2 * h_bar) * pi_over_180) + \
          (T)0.32 * COS(((T)3 * h_bar + (T)6) * pi_over_180) - (T)0.20 * COS(((T)4 * h_bar - (T)63) * pi_over_180); \
    T s_l = (T)1 + ((T)0.015 * lb50) / SQRT((T)20 + lb50);                                              \
    T s_c = (T)1 + (T)0.045 * c_bar_p;                                                                  \
    T s_h = (T)1 + (T)0.015 * c_bar_p * t;                                                              \
    T hb = (h_bar - (T)275) / (T)25;                                                                    \
    T delta_theta = (T)30 * EXP(-(hb * hb));                                                            \
    T cbp2 = c_bar_p * c_bar_p;                                                                         \
    T cbp7 = cbp2 * cbp2 * cbp2 * c_bar_p;                                                              \
    T r_c = (T)2 * SQRT(cbp7 / (cbp7 + twenty_five_pow_seven));                                         \
    T r_t = -r_c * SIN((T)2 * delta_theta * pi_over_180);                                               \
    T delta_c_prime = c2p - c1p;                                                                        \
    T tl = delta_l_prime / s_l;                                                                         \
    T tc = delta_c_prime / s_c;                                                                         \
    T th = delta_big_h / s_h;                                                                           \
    return SQRT(tl * tl + tc * tc + th * th + r_t * tc * th);

float ora_ciede2000_f32(const float lab1[3], const float lab2[3]) {
    CIEDE2000_BODY(float, sqrtf, atan2f, sinf, cosf, expf, fabsf)
}

double ora_ciede2000_f64(const double lab1[3], const double lab2[3]) {
    CIEDE2000_BODY(double, sqrt, atan2, sin, cos, exp, fabs)
}

/* ============================================================================================ */
/* colour distances (lib.rs:1080-1100)                                                            */
/* ============================================================================================ */
double ora_color_distance_red_mean(const uint8_t c1[3], const uint8_t c2[3]) { /* lib.rs:1080-1088 */
    double red_mean = ((double)c1[0] + (double)c2[0]) / 2.0; /* f64::midpoint, exact for these operands */
    double r = (double)c1[0] - (double)c2[0];
    double g = (double)c1[1] - (double)c2[1];
    double b = (double)c1[2] - (double)c2[2];
    return sqrt((((512.0 + red_mean) * r * r) / 256.0) + 4.0 * g * g + (((767.0 - red_mean) * b * b) / 256.0));
}

double ora_color_distance_cielab(const uint8_t c1[3], const uint8_t c2[3]) { /* lib.rs:1090-1100 (memo dropped) */
    float l1[3], l2[3];
    ora_srgb8_to_lab_f32(c1[0], c1[1], c1[2], l1);
    ora_srgb8_to_lab_f32(c2[0], c2[1], c2[2], l2);
    return (double)ora_ciede2000_f32(l1, l2);
}

void ora_new_nes_only(const uint8_t c5[3], int cielab, uint8_t out[3]) { /* lib.rs:640-660 */
    uint8_t color[4], cand[4];
    ora_snes_as_rgba(c5, color);
    int best = 0;
    double best_error = 1.7976931348623157e308; /* f64::MAX */
    for (int index = 0; index < ORA_NES_COLOR_COUNT; index++) {
        ora_snes_as_rgba(NES_TABLE[index], cand);
        double error = cielab ? ora_color_distance_cielab(color, cand) : ora_color_distance_red_mean(color, cand);
        if (error < best_error) {
            best = index;
            best_error = error;
        }
    }
    memcpy(out, NES_TABLE[best], 3);
}

/* Palette::get_closest_color_index (lib.rs:762-795) over an explicit colour list. */
int ora_closest_color_index(const uint8_t *colors5, int n, const double target[3], int cielab) {
    int best_index = 0;
    double best_error = 1.7976931348623157e308;
    uint8_t t8[3];
    for (int i = 0; i < 3; i++) { /* lib.rs:773-778: clamp, round half away from zero, as u8 */
        double v = target[i];
        v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
        v = round(v);
        t8[i] = (v != v) ? 0 : (uint8_t)v;
    }
    for (int index = 0; index < n; index++) {
        uint8_t color[4];
        ora_snes_as_rgba(colors5 + 3 * index, color);
        double error = cielab ? ora_color_distance_cielab(color, t8) : ora_color_distance_red_mean(color, t8);
        if (error < best_error) {
            best_error = error;
            best_index = index;
        }
    }
    return best_index;
}

/* ============================================================================================ */
/* cogset 0.2.0 Kmeans restated: Lloyd, first-k initial centres, tol 1e-6 on the summed squared     */
/* distances, <=100 iterations, empty cluster -> NaN centre (0 * inf).                              */
/* ============================================================================================ */
static void km_update_assignments(const double *pts, int n, int k, const double *centres, int *assign, int *counts,
                                  double *costs) {
    for (int c = 0; c < k; c++) counts[c] = 0;
    for (int i = 0; i < n; i++) {
        int min_index = 0;
        double min_dist = INFINITY;
        for (int c = 0; c < k; c++) {
            double d0 = pts[3 * i] - centres[3 * c];
            double d1 = pts[3 * i + 1] - centres[3 * c + 1];
            double d2 = pts[3 * i + 2] - centres[3 * c + 2];
            double e = (d0 * d0 + d1 * d1) + d2 * d2;
            if (e < min_dist) {
                min_index = c;
                min_dist = e;
            }
        }
        assign[i] = min_index;
        costs[i] = min_dist;
        counts[min_index]++;
    }
}

static void km_update_centres(const double *pts, int n, int k, const int *assign, const int *counts, double *centres) {
    for (int c = 0; c < 3 * k; c++) centres[c] = 0.0;
    for (int i = 0; i < n; i++) {
        double *c = centres + 3 * assign[i];
        c[0] += pts[3 * i];
        c[1] += pts[3 * i + 1];
        c[2] += pts[3 * i + 2];
    }
    for (int c = 0; c < k; c++) {
        double scale = 1.0 / (double)counts[c];
        centres[3 * c] *= scale;
        centres[3 * c + 1] *= scale;
        centres[3 * c + 2] *= scale;
    }
}

int ora_kmeans(const double *pts, int n, int k, double *centres, int *assign) {
    if (!(2 <= k && k < n)) return -1; /* cogset: assert!(2 <= k && k < data.len()) */
    int *counts = (int *)malloc(sizeof(int) * k);
    double *costs = (double *)malloc(sizeof(double) * n);
    memcpy(centres, pts, sizeof(double) * 3 * k);
    km_update_assignments(pts, n, k, centres, assign, counts, costs);
    double objective = 0.0;
    for (int i = 0; i < n; i++) objective += costs[i];
    int iter = 0;
    while (iter < ORA_KMEANS_MAX_ITER) {
        km_update_centres(pts, n, k, assign, counts, centres);
        km_update_assignments(pts, n, k, centres, assign, counts, costs);
        double new_objective = 0.0;
        for (int i = 0; i < n; i++) new_objective += costs[i];
        if (fabs(new_objective - objective) < ORA_KMEANS_TOL) break;
        objective = new_objective;
        iter++;
    }
    free(counts);
    free(costs);
    return iter;
}

/* ============================================================================================ */
/* tile assignment and initial palettes (lib.rs:79-189, 330-415)                                  */
/* ============================================================================================ */
static uint8_t f64_round_as_u8(double v) { /* `(v).round() as u8`: saturating, NaN -> 0 */
    v = round(v);
    if (!(v > 0.0)) return 0;
    if (v > 255.0) return 255;
    return (uint8_t)v;
}

/* centre (f64 triple) -> SnesColor; shared tail of lib.rs:140-171 and 369-401 */
static void centre_to_color(const ora_image *im, const double v[3], uint8_t out[3]) {
    uint8_t c5[3];
    if (im->perceptual_palettes) {
        uint8_t rgb[3];
        ora_lab_f64_to_srgb8(v, rgb);
        c5[0] = rgb[0] / 8;
        c5[1] = rgb[1] / 8;
        c5[2] = rgb[2] / 8;
    } else {
        c5[0] = f64_round_as_u8(v[0] / 8.0);
        c5[1] = f64_round_as_u8(v[1] / 8.0);
        c5[2] = f64_round_as_u8(v[2] / 8.0);
    }
    if (im->nes) ora_new_nes_only(c5, im->perceptual_palettes, out);
    else memcpy(out, c5, 3);
}

static int recalculate_palette(ora_image *im, int palette) { /* lib.rs:330-405 */
    double *pixels = (double *)malloc(sizeof(double) * 3 * ORA_NPIX);
    int n = 0;
    const int wt = im->width / 8;
    for (int tile = 0; tile < ORA_NTILES; tile++) {
        if (im->tile_palettes[tile] != palette) continue;
        int tile_x = tile % wt, tile_y = tile / wt;
        for (int x = 0; x < 8; x++)
            for (int y = 0; y < 8; y++) {
                const uint8_t *c = orig_px(im, tile_x * 8 + x, tile_y * 8 + y);
                if (c[3] == 0) continue;
                if (im->perceptual_palettes) {
                    float lab[3];
                    ora_srgb8_to_lab_f32(c[0], c[1], c[2], lab);
                    pixels[3 * n] = (double)lab[0];
                    pixels[3 * n + 1] = (double)lab[1];
                    pixels[3 * n + 2] = (double)lab[2];
                } else {
                    pixels[3 * n] = (double)c[0];
                    pixels[3 * n + 1] = (double)c[1];
                    pixels[3 * n + 2] = (double)c[2];
                }
                n++;
            }
    }
    const int k = im->sub_size;
    double *centres = (double *)malloc(sizeof(double) * 3 * (k > 0 ? k : 1));
    int *assign = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    int rc = ora_kmeans(pixels, n, k, centres, assign);
    if (rc >= 0)
        for (int index = 0; index < k; index++) centre_to_color(im, centres + 3 * index, im->colors + 3 * (palette * k + index));
    free(pixels);
    free(centres);
    free(assign);
    return rc < 0 ? -1 : 0;
}

int ora_recalculate_palettes(ora_image *im) { /* lib.rs:407-415 */
    for (int p = 0; p < im->sub_count; p++)
        if (recalculate_palette(im, p) != 0) return -1;
    ora_optimize(im);
    return 0;
}

int ora_initialize_tiles(ora_image *im) { /* lib.rs:79-189 */
    if (im->sub_count == 1) {
        if (recalculate_palette(im, 0) != 0) return -1;
        ora_optimize(im);
        return 0;
    }
    const int wt = im->width / 8, ht = im->height / 8;
    double *means = (double *)malloc(sizeof(double) * 3 * ORA_NTILES);
    int *map = (int *)malloc(sizeof(int) * ORA_NTILES);
    int n = 0;
    for (int tile_x = 0; tile_x < wt; tile_x++)
        for (int tile_y = 0; tile_y < ht; tile_y++) {
            float sum[3] = {0.0f, 0.0f, 0.0f};
            int count = 0;
            int index = tile_y * wt + tile_x;
            for (int x = 0; x < 8; x++)
                for (int y = 0; y < 8; y++) {
                    const uint8_t *c = orig_px(im, tile_x * 8 + x, tile_y * 8 + y);
                    if (c[3] == 0) continue;
                    if (im->perceptual_palettes) {
                        float lab[3];
                        ora_srgb8_to_lab_f32(c[0], c[1], c[2], lab);
                        sum[0] += lab[0];
                        sum[1] += lab[1];
                        sum[2] += lab[2];
                    } else {
                        sum[0] += (float)c[0];
                        sum[1] += (float)c[1];
                        sum[2] += (float)c[2];
                    }
                    count++;
                }
            if ((sum[0] + sum[1]) + sum[2] > 0.0f) { /* lib.rs:118 */
                means[3 * n] = (double)sum[0] / (double)count;
                means[3 * n + 1] = (double)sum[1] / (double)count;
                means[3 * n + 2] = (double)sum[2] / (double)count;
                map[n] = index;
                n++;
            }
        }
    const int k = im->sub_count;
    double *centres = (double *)malloc(sizeof(double) * 3 * k);
    int *assign = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    int rc = ora_kmeans(means, n, k, centres, assign);
    if (rc >= 0) {
        for (int i = 0; i < n; i++) im->tile_palettes[map[i]] = (uint8_t)assign[i]; /* lib.rs:133-138 */
        for (int index = 0; index < k; index++) {
            uint8_t color[3];
            centre_to_color(im, centres + 3 * index, color);
            for (int i = 0; i < im->sub_size; i++) memcpy(im->colors + 3 * (index * im->sub_size + i), color, 3); /* 181-183 */
        }
        ora_optimize(im);
    }
    free(means);
    free(map);
    free(centres);
    free(assign);
    return rc < 0 ? -1 : 0;
}

/* ============================================================================================ */
/* optimize (lib.rs:417-501)                                                                      */
/* ============================================================================================ */
void ora_optimize(ora_image *im) {
    const int W = im->width, H = im->height;
    double dither_weights[4] = {0.0, 0.0, 0.0, 0.0};
    if (im->dither) {
        dither_weights[0] = 7.0 / 16.0;
        dither_weights[1] = 3.0 / 16.0;
        dither_weights[2] = 5.0 / 16.0;
        dither_weights[3] = 1.0 / 16.0;
    }
    const double error_multiplier = 0.8;
    double *error = (double *)calloc((size_t)W * H * 3, sizeof(double));
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int pixel_index = y * W + x;
            const uint8_t *oc = orig_px(im, x, y);
            const int palette = im->tile_palettes[(x / 8) + (y / 8) * (W / 8)]; /* lib.rs:417-423 */
            double target[3] = {(double)oc[0] + error[3 * pixel_index], (double)oc[1] + error[3 * pixel_index + 1],
                                (double)oc[2] + error[3 * pixel_index + 2]};
            const uint8_t *sub = im->colors + 3 * (palette * im->sub_size);
            int color_index = ora_closest_color_index(sub, im->sub_size, target, im->perceptual_palettes);
            im->palette_map[pixel_index] = oc[3] > 0 ? (uint8_t)color_index : 0;
            uint8_t nc[4];
            ora_snes_as_rgba(sub + 3 * color_index, nc);
            double pixel_error[3];
            if (oc[3] > 0) {
                pixel_error[0] = target[0] - (double)nc[0];
                pixel_error[1] = target[1] - (double)nc[1];
                pixel_error[2] = target[2] - (double)nc[2];
            } else {
                pixel_error[0] = error[3 * pixel_index];
                pixel_error[1] = error[3 * pixel_index + 1];
                pixel_error[2] = error[3 * pixel_index + 2];
            }
            for (int i = 0; i < 3; i++) {
                double value = pixel_error[i];
                if (x + 1 < W) error[3 * (pixel_index + 1) + i] += value * error_multiplier * dither_weights[0];
                if (y + 1 < H) {
                    if (x > 0) error[3 * (pixel_index + W - 1) + i] += value * error_multiplier * dither_weights[1];
                    error[3 * (pixel_index + W) + i] += value * error_multiplier * dither_weights[2];
                    if (x + 1 < W) error[3 * (pixel_index + W + 1) + i] += value * error_multiplier * dither_weights[3];
                }
            }
        }
    free(error);
}

void ora_as_rgba(const ora_image *im, uint8_t *out) { /* lib.rs:550-577 */
    memset(out, 0, (size_t)ORA_NPIX * 4);
    for (int y = 0; y < im->height; y++)
        for (int x = 0; x < im->width; x++) {
            int palette_index = im->tile_palettes[(y / 8) * 32 + (x / 8)];
            int color_index = palette_index * im->sub_size + im->palette_map[y * im->width + x];
            if (orig_px(im, x, y)[3] > 0) ora_snes_as_rgba(im->colors + 3 * color_index, out + 4 * (y * im->width + x));
        }
}

void ora_as_json_arrays(const ora_image *im, uint16_t *palette16, uint8_t *tiles, uint8_t *tile_palettes) { /* 579-625 */
    for (int p = 0; p < im->sub_count; p++)
        for (int i = 0; i < 16; i++) {
            uint16_t v = 0;
            if (i != 0 && i <= im->sub_size) v = ora_snes_as_u16(im->colors + 3 * (p * im->sub_size + i - 1));
            palette16[p * 16 + i] = v;
        }
    const int wt = im->width / 8, ht = im->height / 8;
    for (int tile_y = 0; tile_y < ht; tile_y++)
        for (int tile_x = 0; tile_x < wt; tile_x++) {
            int tile_index = tile_y * wt + tile_x;
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    int index = (tile_y * 8 + y) * im->width + (tile_x * 8 + x);
                    tiles[tile_index * 64 + y * 8 + x] =
                        orig_px(im, tile_x * 8 + x, tile_y * 8 + y)[3] == 0 ? 0 : (uint8_t)(im->palette_map[index] + 1);
                }
            tile_palettes[tile_index] = im->tile_palettes[tile_index];
        }
}

/* ============================================================================================ */
/* ssimulacra2 0.5.1 / yuvxyb 0.4.2 restated                                                      */
/* ============================================================================================ */
static float srgb_eotf_formula(float x) { /* yuvxyb transfer: sRGB -> linear (zimg constants) */
    x = x > 0.0f ? x : 0.0f;
    if (x < 12.92f * ORA_SRGB_BETA) return x / 12.92f;
    return powf((x + (ORA_SRGB_ALPHA - 1.0f)) / ORA_SRGB_ALPHA, 2.4f);
}
float ora_srgb_eotf(float x) { return srgb_eotf_formula(x); }

float ora_cbrtf(float x) { /* FreeBSD msun s_cbrtf.c as ported by yuvxyb-math */
    const uint32_t B1 = 709958130u, B2 = 642849266u;
    union {
        float f;
        uint32_t i;
    } u = {x};
    uint32_t hx = u.i & 0x7fffffffu;
    uint32_t sign = u.i & 0x80000000u;
    if (hx >= 0x7f800000u) return x + x;
    if (hx < 0x00800000u) {
        if (hx == 0) return x;
        u.f = x * 0x1p24f;
        hx = u.i & 0x7fffffffu;
        hx = hx / 3 + B2;
    } else {
        hx = hx / 3 + B1;
    }
    u.i = sign | hx;
    double t = (double)u.f;
    double r = t * t * t;
    t = t * ((double)x + (double)x + r) / ((double)x + r + r);
    r = t * t * t;
    t = t * ((double)x + (double)x + r) / ((double)x + r + r);
    return (float)t;
}

void ora_linear_rgb_to_xyb(const float rgb[3], float xyb[3]) { /* yuvxyb linear_rgb_to_xyb */
    float m0 = fmaf(ORA_K_M00, rgb[0], fmaf(ORA_K_M01, rgb[1], fmaf(ORA_K_M02, rgb[2], ORA_K_B0)));
    float m1 = fmaf(ORA_K_M10, rgb[0], fmaf(ORA_K_M11, rgb[1], fmaf(ORA_K_M12, rgb[2], ORA_K_B0)));
    float m2 = fmaf(ORA_K_M20, rgb[0], fmaf(ORA_K_M21, rgb[1], fmaf(ORA_K_M22, rgb[2], ORA_K_B0)));
    m0 = m0 < 0.0f ? 0.0f : m0;
    m1 = m1 < 0.0f ? 0.0f : m1;
    m2 = m2 < 0.0f ? 0.0f : m2;
    m0 = ora_cbrtf(m0) + ORA_NEG_CBRT_BIAS;
    m1 = ora_cbrtf(m1) + ORA_NEG_CBRT_BIAS;
    m2 = ora_cbrtf(m2) + ORA_NEG_CBRT_BIAS;
    xyb[0] = 0.5f * (m0 - m1);
    xyb[1] = 0.5f * (m0 + m1);
    xyb[2] = m2;
}

static inline void make_positive_xyb(float p[3]) { /* ssimulacra2 make_positive_xyb */
    p[2] = (p[2] - p[1]) + 0.55f;
    p[0] = fmaf(p[0], 14.0f, 0.42f);
    p[1] = p[1] + 0.01f;
}

/* libjxl CreateRecursiveGaussian(sigma): Charalampidis 2016 3-term recursive Gaussian. */
static void inv3x3(double m[9]) {
    double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    double det = a * A + b * B + c * C;
    double inv = 1.0 / det;
    m[0] = A * inv;
    m[1] = -(b * i - c * h) * inv;
    m[2] = (b * f - c * e) * inv;
    m[3] = B * inv;
    m[4] = (a * i - c * g) * inv;
    m[5] = -(a * f - c * d) * inv;
    m[6] = C * inv;
    m[7] = -(a * h - b * g) * inv;
    m[8] = (a * e - b * d) * inv;
}

void ora_gaussian_coeffs(float n2_out[3], float d1_out[3], int *radius_out) {
    const double sigma = ORA_BLUR_SIGMA;
    const double kPi = 3.141592653589793238;
    const double radius = round(3.2795 * sigma + 0.2546);
    const double pi_div_2r = kPi / (2.0 * radius);
    const double omega[3] = {pi_div_2r, 3.0 * pi_div_2r, 5.0 * pi_div_2r};
    const double p_1 = +1.0 / tan(0.5 * omega[0]);
    const double p_3 = -1.0 / tan(0.5 * omega[1]);
    const double p_5 = +1.0 / tan(0.5 * omega[2]);
    const double r_1 = +p_1 * p_1 / sin(omega[0]);
    const double r_3 = -p_3 * p_3 / sin(omega[1]);
    const double r_5 = +p_5 * p_5 / sin(omega[2]);
    const double neg_half_sigma2 = -0.5 * sigma * sigma;
    const double recip_radius = 1.0 / radius;
    double rho[3];
    for (int i = 0; i < 3; i++) rho[i] = exp(neg_half_sigma2 * omega[i] * omega[i]) * recip_radius;
    const double D_13 = p_1 * r_3 - r_1 * p_3;
    const double D_35 = p_3 * r_5 - r_3 * p_5;
    const double D_51 = p_5 * r_1 - r_5 * p_1;
    const double recip_d13 = 1.0 / D_13;
    const double zeta_15 = D_35 * recip_d13;
    const double zeta_35 = D_51 * recip_d13;
    double A[9] = {p_1, p_3, p_5, r_1, r_3, r_5, zeta_15, zeta_35, 1.0};
    inv3x3(A);
    const double gamma[3] = {1.0, radius * radius - sigma * sigma, zeta_15 * rho[0] + zeta_35 * rho[1] + rho[2]};
    double beta[3];
    for (int i = 0; i < 3; i++) beta[i] = A[3 * i] * gamma[0] + A[3 * i + 1] * gamma[1] + A[3 * i + 2] * gamma[2];
    for (int i = 0; i < 3; i++) {
        n2_out[i] = (float)(-beta[i] * cos(omega[i] * (radius + 1.0)));
        d1_out[i] = (float)(-2.0 * cos(omega[i]));
    }
    *radius_out = (int)radius;
}

typedef struct {
    float n2[3], d1[3];
    int radius;
    int ready;
} gauss_t;
static gauss_t G;
static void gauss_init(void) {
    if (!G.ready) {
        ora_gaussian_coeffs(G.n2, G.d1, &G.radius);
        G.ready = 1;
    }
}

/* horizontal pass of one row: out = n2*sum; out = fma(-1, prev2, out); out = fma(-d1, prev, out) */
static void blur_h_row(const float *in, float *out, int width) {
    const int N = G.radius;
    const float mul_in_1 = G.n2[0], mul_in_3 = G.n2[1], mul_in_5 = G.n2[2];
    const float mul_prev_1 = -G.d1[0], mul_prev_3 = -G.d1[1], mul_prev_5 = -G.d1[2];
    const float mul_prev2 = -1.0f;
    float prev_1 = 0, prev_3 = 0, prev_5 = 0, prev2_1 = 0, prev2_3 = 0, prev2_5 = 0;
    for (int n = -N + 1; n < width; n++) {
        int left = n - N - 1, right = n + N - 1;
        float left_val = left >= 0 ? in[left] : 0.0f;
        float right_val = right < width ? in[right] : 0.0f;
        float sum = left_val + right_val;
        float out_1 = sum * mul_in_1;
        float out_3 = sum * mul_in_3;
        float out_5 = sum * mul_in_5;
        out_1 = fmaf(mul_prev2, prev2_1, out_1);
        out_3 = fmaf(mul_prev2, prev2_3, out_3);
        out_5 = fmaf(mul_prev2, prev2_5, out_5);
        prev2_1 = prev_1;
        prev2_3 = prev_3;
        prev2_5 = prev_5;
        out_1 = fmaf(mul_prev_1, prev_1, out_1);
        out_3 = fmaf(mul_prev_3, prev_3, out_3);
        out_5 = fmaf(mul_prev_5, prev_5, out_5);
        prev_1 = out_1;
        prev_3 = out_3;
        prev_5 = out_5;
        if (n >= 0) out[n] = (out_1 + out_3) + out_5;
    }
}

/* vertical pass of one column: t = fma(prev, d1, prev2); out = fma(sum, n2, -t) */
static void blur_v_col(const float *in, float *out, int width, int height, int col) {
    const int N = G.radius;
    float prev_1 = 0, prev_3 = 0, prev_5 = 0, prev2_1 = 0, prev2_3 = 0, prev2_5 = 0;
    for (int n = -N + 1; n < height; n++) {
        int top = n - N - 1, bottom = n + N - 1;
        float top_val = top >= 0 ? in[(size_t)top * width + col] : 0.0f;
        float bottom_val = bottom < height ? in[(size_t)bottom * width + col] : 0.0f;
        float sum = top_val + bottom_val;
        float t1 = fmaf(prev_1, G.d1[0], prev2_1);
        float t3 = fmaf(prev_3, G.d1[1], prev2_3);
        float t5 = fmaf(prev_5, G.d1[2], prev2_5);
        float out_1 = fmaf(sum, G.n2[0], -t1);
        float out_3 = fmaf(sum, G.n2[1], -t3);
        float out_5 = fmaf(sum, G.n2[2], -t5);
        prev2_1 = prev_1;
        prev2_3 = prev_3;
        prev2_5 = prev_5;
        prev_1 = out_1;
        prev_3 = out_3;
        prev_5 = out_5;
        if (n >= 0) out[(size_t)n * width + col] = (out_1 + out_3) + out_5;
    }
}

static void blur_plane_tmp(const float *in, float *out, float *tmp, int w, int h) {
    gauss_init();
    for (int y = 0; y < h; y++) blur_h_row(in + (size_t)y * w, tmp + (size_t)y * w, w);
    for (int x = 0; x < w; x++) blur_v_col(tmp, out, w, h, x);
}

void ora_blur_plane(const float *in, float *out, int w, int h) {
    float *tmp = (float *)malloc(sizeof(float) * (size_t)w * h);
    blur_plane_tmp(in, out, tmp, w, h);
    free(tmp);
}

/* interleaved linear RGB image */
typedef struct {
    int w, h;
    float *d; /* w*h*3 */
} linrgb_t;

static linrgb_t lin_from_rgba8(const uint8_t *rgba, int w, int h) { /* lib.rs:506-516 + yuvxyb Rgb->LinearRgb */
    luts_init();
    linrgb_t o = {w, h, (float *)malloc(sizeof(float) * 3 * (size_t)w * h)};
    for (size_t i = 0; i < (size_t)w * h; i++)
        for (int c = 0; c < 3; c++) o.d[3 * i + c] = g_lut_yuvxyb[rgba[4 * i + c]]; /* ora_srgb_eotf(v / 255) unless a table was injected */
    return o;
}

static linrgb_t downscale_by_2(const linrgb_t *in) { /* ssimulacra2 downscale_by_2 */
    const int in_w = in->w, in_h = in->h;
    const int out_w = (in_w + 1) / 2, out_h = (in_h + 1) / 2;
    linrgb_t o = {out_w, out_h, (float *)malloc(sizeof(float) * 3 * (size_t)out_w * out_h)};
    const float normalize = 1.0f / 4.0f;
    for (int oy = 0; oy < out_h; oy++)
        for (int ox = 0; ox < out_w; ox++)
            for (int c = 0; c < 3; c++) {
                float sum = 0.0f;
                for (int iy = 0; iy < 2; iy++)
                    for (int ix = 0; ix < 2; ix++) {
                        int x = ox * 2 + ix, y = oy * 2 + iy;
                        if (x > in_w - 1) x = in_w - 1;
                        if (y > in_h - 1) y = in_h - 1;
                        sum += in->d[3 * ((size_t)y * in_w + x) + c];
                    }
                o.d[3 * ((size_t)oy * out_w + ox) + c] = sum * normalize;
            }
    return o;
}

/* linear RGB -> positive XYB, planar: p[c][y*w+x] */
static void lin_to_xyb_planar(const linrgb_t *in, float *planes) {
    const size_t n = (size_t)in->w * in->h;
    for (size_t i = 0; i < n; i++) {
        float xyb[3];
        ora_linear_rgb_to_xyb(in->d + 3 * i, xyb);
        make_positive_xyb(xyb);
        planes[i] = xyb[0];
        planes[n + i] = xyb[1];
        planes[2 * n + i] = xyb[2];
    }
}

static void ssim_map(int w, int h, const float *m1, const float *m2, const float *s11, const float *s22, const float *s12,
                     double out[6]) {
    const size_t n = (size_t)w * h;
    const double one_per_pixels = 1.0 / (double)n;
    for (int c = 0; c < 3; c++) {
        double sum0 = 0.0, sum1 = 0.0;
        for (size_t i = c * n; i < (c + 1) * n; i++) {
            float mu1 = m1[i], mu2 = m2[i];
            float mu11 = mu1 * mu1, mu22 = mu2 * mu2, mu12 = mu1 * mu2;
            float mu_diff = mu1 - mu2;
            float num_m = fmaf(mu_diff, -mu_diff, 1.0f);
            float num_s = fmaf(2.0f, s12[i] - mu12, ORA_SSIM_C2);
            float denom_s = (s11[i] - mu11) + (s22[i] - mu22) + ORA_SSIM_C2;
            double d = 1.0 - (double)((num_m * num_s) / denom_s);
            d = fmax(d, 0.0);
            sum0 += d;
            double d2 = d * d;
            sum1 += d2 * d2;
        }
        out[c * 2] = one_per_pixels * sum0;
        out[c * 2 + 1] = sqrt(sqrt(one_per_pixels * sum1));
    }
}

static void edge_diff_map(int w, int h, const float *i1, const float *m1, const float *i2, const float *m2, double out[12]) {
    const size_t n = (size_t)w * h;
    const double one_per_pixels = 1.0 / (double)n;
    for (int c = 0; c < 3; c++) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (size_t i = c * n; i < (c + 1) * n; i++) {
            double d1 = (1.0 + (double)fabsf(i2[i] - m2[i])) / (1.0 + (double)fabsf(i1[i] - m1[i])) - 1.0;
            double artifact = fmax(d1, 0.0);
            s0 += artifact;
            double a2 = artifact * artifact;
            s1 += a2 * a2;
            double detail_lost = fmax(-d1, 0.0);
            s2 += detail_lost;
            double l2 = detail_lost * detail_lost;
            s3 += l2 * l2;
        }
        out[c * 4] = one_per_pixels * s0;
        out[c * 4 + 1] = sqrt(sqrt(one_per_pixels * s1));
        out[c * 4 + 2] = one_per_pixels * s2;
        out[c * 4 + 3] = sqrt(sqrt(one_per_pixels * s3));
    }
}

static double msssim_score(const double *avg /* [nscales][18] */, int nscales) { /* Msssim::score */
    double ssim = 0.0;
    int i = 0;
    for (int c = 0; c < 3; c++)
        for (int s = 0; s < nscales; s++) {
            const double *a = avg + 18 * s;
            for (int n = 0; n < 2; n++) {
                ssim = fma(ORA_SSIM2_WEIGHT[i++], fabs(a[c * 2 + n]), ssim);
                ssim = fma(ORA_SSIM2_WEIGHT[i++], fabs(a[6 + c * 4 + n]), ssim);
                ssim = fma(ORA_SSIM2_WEIGHT[i++], fabs(a[6 + c * 4 + n + 2]), ssim);
            }
        }
    ssim *= ORA_POOL_SCALE;
    ssim = fma(ORA_POOL_C3 * ssim * ssim, ssim, fma(ORA_POOL_C1, ssim, ORA_POOL_C2 * ssim * ssim));
    if (ssim > 0.0) ssim = fma(pow(ssim, ORA_POOL_EXP), -10.0, 100.0);
    else ssim = 100.0;
    return ssim;
}

static double ssimulacra2_lin(linrgb_t img1, linrgb_t img2, double *avg_out) { /* consumes img1/img2 */
    int w = img1.w, h = img1.h;
    const size_t n0 = (size_t)w * h;
    float *buf = (float *)malloc(sizeof(float) * 3 * n0 * 9);
    float *p1 = buf, *p2 = buf + 3 * n0, *mul = buf + 6 * n0, *s11 = buf + 9 * n0, *s22 = buf + 12 * n0,
          *s12 = buf + 15 * n0, *mu1 = buf + 18 * n0, *mu2 = buf + 21 * n0, *tmp = buf + 24 * n0;
    double avg[ORA_NUM_SCALES * 18];
    int nscales = 0;
    for (int scale = 0; scale < ORA_NUM_SCALES; scale++) {
        if (w < 8 || h < 8) break;
        if (scale > 0) {
            linrgb_t a = downscale_by_2(&img1), b = downscale_by_2(&img2);
            free(img1.d);
            free(img2.d);
            img1 = a;
            img2 = b;
            w = img1.w;
            h = img1.h;
        }
        const size_t n = (size_t)w * h;
        lin_to_xyb_planar(&img1, p1);
        lin_to_xyb_planar(&img2, p2);
        for (size_t i = 0; i < 3 * n; i++) mul[i] = p1[i] * p1[i];
        for (int c = 0; c < 3; c++) blur_plane_tmp(mul + c * n, s11 + c * n, tmp, w, h);
        for (size_t i = 0; i < 3 * n; i++) mul[i] = p2[i] * p2[i];
        for (int c = 0; c < 3; c++) blur_plane_tmp(mul + c * n, s22 + c * n, tmp, w, h);
        for (size_t i = 0; i < 3 * n; i++) mul[i] = p1[i] * p2[i];
        for (int c = 0; c < 3; c++) blur_plane_tmp(mul + c * n, s12 + c * n, tmp, w, h);
        for (int c = 0; c < 3; c++) blur_plane_tmp(p1 + c * n, mu1 + c * n, tmp, w, h);
        for (int c = 0; c < 3; c++) blur_plane_tmp(p2 + c * n, mu2 + c * n, tmp, w, h);
        ssim_map(w, h, mu1, mu2, s11, s22, s12, avg + 18 * nscales);
        edge_diff_map(w, h, p1, mu1, p2, mu2, avg + 18 * nscales + 6);
        nscales++;
    }
    free(buf);
    free(img1.d);
    free(img2.d);
    if (avg_out) memcpy(avg_out, avg, sizeof(double) * 18 * nscales);
    return msssim_score(avg, nscales);
}

double ora_ssimulacra2_rgba8(const uint8_t *src_rgba, const uint8_t *dst_rgba, int w, int h, double *avg) {
    if (w < 8 || h < 8) return NAN;
    return ssimulacra2_lin(lin_from_rgba8(src_rgba, w, h), lin_from_rgba8(dst_rgba, w, h), avg);
}

void ora_xyb_pyramid_rgba8(const uint8_t *rgba, int w, int h, float *out) {
    linrgb_t img = lin_from_rgba8(rgba, w, h);
    for (int scale = 0; scale < ORA_NUM_SCALES; scale++) {
        if (img.w < 8 || img.h < 8) break;
        if (scale > 0) {
            linrgb_t a = downscale_by_2(&img);
            free(img.d);
            img = a;
        }
        lin_to_xyb_planar(&img, out);
        out += 3 * (size_t)img.w * img.h;
    }
    free(img.d);
}

void ora_source_planes_rgba8(const uint8_t *rgba, int w, int h, float *mu1, float *s11) {
    float *xyb = (float *)malloc(sizeof(float) * 3 * ORA_TOTAL_SCALE_PIXELS);
    float *mul = (float *)malloc(sizeof(float) * (size_t)w * h);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)w * h);
    ora_xyb_pyramid_rgba8(rgba, w, h, xyb);
    size_t off = 0;
    for (int scale = 0; scale < ORA_NUM_SCALES; scale++) {
        const size_t n = (size_t)w * h;
        for (int c = 0; c < 3; c++) {
            const float *p = xyb + off + c * n;
            blur_plane_tmp(p, mu1 + off + c * n, tmp, w, h);
            for (size_t i = 0; i < n; i++) mul[i] = p[i] * p[i];
            blur_plane_tmp(mul, s11 + off + c * n, tmp, w, h);
        }
        off += 3 * n;
        w = (w + 1) / 2;
        h = (h + 1) / 2;
    }
    free(xyb);
    free(mul);
    free(tmp);
}

double ora_error(const ora_image *im) { /* lib.rs:503-548 */
    uint8_t *rgba = (uint8_t *)malloc((size_t)ORA_NPIX * 4);
    ora_as_rgba(im, rgba);
    double score = ora_ssimulacra2_rgba8(im->original, rgba, im->width, im->height, NULL);
    free(rgba);
    return 100.0 - score;
}

/* ============================================================================================ */
/* palette-entry optimisers (lib.rs:191-328)                                                      */
/* ============================================================================================ */
void ora_eval_candidates(ora_image *im, int palette, int index, const uint8_t *cand, int ncand, double *scores,
                         uint8_t *maps) {
    uint8_t *slot = im->colors + 3 * (palette * im->sub_size + index);
    uint8_t saved[3];
    memcpy(saved, slot, 3);
    for (int k = 0; k < ncand; k++) {
        memcpy(slot, cand + 3 * k, 3);
        ora_optimize(im);
        scores[k] = ora_error(im);
        if (maps) memcpy(maps + (size_t)k * ORA_NPIX, im->palette_map, ORA_NPIX);
    }
    memcpy(slot, saved, 3);
    ora_optimize(im);
}

int ora_optimize_palette_entry_random(ora_image *im, int palette, int index, const uint8_t *cand, int ncand) {
    uint8_t *slot = im->colors + 3 * (palette * im->sub_size + index);
    uint8_t best_color[3];
    memcpy(best_color, slot, 3);
    double best_error = ora_error(im); /* lib.rs:199 */
    for (int k = 0; k < ncand; k++) {   /* lib.rs:205-220 */
        memcpy(slot, cand + 3 * k, 3);
        ora_optimize(im);
        double error = ora_error(im);
        if (error < best_error) {
            best_error = error;
            memcpy(best_color, cand + 3 * k, 3);
        }
    }
    memcpy(slot, best_color, 3); /* lib.rs:236-237 */
    ora_optimize(im);
    return 0;
}

int ora_optimize_palette_entry_nes(ora_image *im, int palette, int index) { /* lib.rs:242-284 */
    uint8_t *slot = im->colors + 3 * (palette * im->sub_size + index);
    int best_index = 0;
    double best_error = 1.7976931348623157e308;
    for (int nes_index = 0; nes_index < ORA_NES_COLOR_COUNT; nes_index++) {
        memcpy(slot, NES_TABLE[nes_index], 3);
        ora_optimize(im);
        double error = ora_error(im);
        if (error < best_error) {
            best_error = error;
            best_index = nes_index;
        }
    }
    memcpy(slot, NES_TABLE[best_index], 3);
    ora_optimize(im);
    return 0;
}

int ora_optimize_palette_entry_channel(ora_image *im, int palette, int index, int channel) { /* lib.rs:286-328 */
    uint8_t *slot = im->colors + 3 * (palette * im->sub_size + index);
    uint8_t best_value = slot[channel];
    double best_error = ora_error(im);
    for (int value = 0; value < 32; value++) {
        slot[channel] = (uint8_t)value;
        ora_optimize(im);
        double error = ora_error(im);
        if (error < best_error) {
            best_error = error;
            best_value = (uint8_t)value;
        }
    }
    slot[channel] = best_value;
    ora_optimize(im);
    return 0;
}
