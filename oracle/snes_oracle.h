/*
 * oracle/snes_oracle.h  --  TEST INFRASTRUCTURE: CPU restatement of the snesimage hot path.
 *
 * Plain-C restatement of /root/reference/src/lib.rs (OptimizedImage, Palette, SnesColor, colour
 * distances) plus the arithmetic of the crates it calls.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (snesimage_b200/) never does.  See constants_unverified.h for what is and is not pinned.
 */
#ifndef SNES_ORACLE_H
#define SNES_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_WIDTH 256
#define ORA_HEIGHT 256
#define ORA_NPIX (ORA_WIDTH * ORA_HEIGHT)
#define ORA_NTILES 1024
#define ORA_NES_COLOR_COUNT 56
#define ORA_TOTAL_SCALE_PIXELS 87360 /* 65536+16384+4096+1024+256+64 */

typedef struct ora_image ora_image; /* struct OptimizedImage, lib.rs:33-43 */

/* OptimizedImage::new, lib.rs:46-65.  Returns NULL unless the image is 256x256. */
ora_image *ora_image_new(const uint8_t *rgba, int width, int height, int sub_count, int sub_size, int dither,
                         int perceptual_palettes, int nes);
void ora_image_free(ora_image *im);

/* lib.rs:79-189.  Returns 0, or -1 where cogset would panic (k < 2 or k >= n). */
int ora_initialize_tiles(ora_image *im);
/* lib.rs:407-415 (and 330-405 per subpalette). */
int ora_recalculate_palettes(ora_image *im);
/* lib.rs:425-501 */
void ora_optimize(ora_image *im);
/* lib.rs:503-548: 100 - SSIMULACRA2(original, as_rgba()).  Recomputes the source side each call,
 * as the reference does. */
double ora_error(const ora_image *im);
/* lib.rs:550-577 */
void ora_as_rgba(const ora_image *im, uint8_t *out_rgba);

/* lib.rs:191-240 with the 64 rand::rng() draws replaced by an explicit list cand[ncand][3]. */
int ora_optimize_palette_entry_random(ora_image *im, int palette, int index, const uint8_t *cand, int ncand);
/* lib.rs:242-284 */
int ora_optimize_palette_entry_nes(ora_image *im, int palette, int index);
/* lib.rs:286-328 */
int ora_optimize_palette_entry_channel(ora_image *im, int palette, int index, int channel);

/* The inner loop of the three optimisers, exposed for differential tests and the CPU baseline:
 * for each candidate colour k: colors[palette*sub_size+index] = cand[k]; optimize(); scores[k]=error().
 * The entry is restored afterwards and optimize() is re-run.  maps (optional) receives the ncand
 * palette_maps (ncand*65536 bytes). */
void ora_eval_candidates(ora_image *im, int palette, int index, const uint8_t *cand, int ncand, double *scores,
                         uint8_t *maps);

/* state accessors */
void ora_get_palette(const ora_image *im, uint8_t *out /* sub_count*sub_size*3, 5-bit values */);
void ora_set_palette(ora_image *im, const uint8_t *in);
void ora_get_tile_palettes(const ora_image *im, uint8_t *out /* 1024 */);
void ora_set_tile_palettes(ora_image *im, const uint8_t *in);
void ora_get_palette_map(const ora_image *im, uint8_t *out /* 65536 */);
void ora_set_palette_map(ora_image *im, const uint8_t *in);
/* lib.rs:579-625 as arrays: palette16[sub_count*16] (u16), tiles[1024*64], tile_palettes[1024] */
void ora_as_json_arrays(const ora_image *im, uint16_t *palette16, uint8_t *tiles, uint8_t *tile_palettes);

/* ---- colour primitives (lib.rs:628-795, 1080-1100) ---------------------------------------- */
void ora_nes_color(int index, uint8_t out[3]);                           /* lib.rs:685-745 */
void ora_snes_as_rgba(const uint8_t c5[3], uint8_t out[4]);              /* lib.rs:662-669 (u8 wrapping) */
uint16_t ora_snes_as_u16(const uint8_t c5[3]);                           /* lib.rs:679-681 */
void ora_new_nes_only(const uint8_t c5[3], int cielab, uint8_t out[3]);  /* lib.rs:640-660 */
double ora_color_distance_red_mean(const uint8_t a[3], const uint8_t b[3]); /* lib.rs:1080-1088 */
double ora_color_distance_cielab(const uint8_t a[3], const uint8_t b[3]);   /* lib.rs:1090-1100 */
/* lib.rs:762-795 on an explicit list of n 5-bit colours */
int ora_closest_color_index(const uint8_t *colors5, int n, const double target[3], int cielab);

/* ---- third-party arithmetic, exposed for known-answer tests ------------------------------- */
void ora_srgb8_to_lab_f32(uint8_t r, uint8_t g, uint8_t b, float out[3]);   /* palette: Srgb<u8>->Lab<D65,f32> */
void ora_lab_f64_to_srgb8(const double lab[3], uint8_t out[3]);             /* palette: Lab<f64>->Srgb<u8> */
float ora_ciede2000_f32(const float lab1[3], const float lab2[3]);          /* palette: Ciede2000 */
double ora_ciede2000_f64(const double lab1[3], const double lab2[3]);       /* same formula in f64 (KAT) */
float ora_srgb_eotf(float v);                                               /* yuvxyb sRGB -> linear */
/* the two 256-entry sRGB -> linear tables the path uses (index = 8-bit value): yuvxyb's and palette's; NULL keeps the built-in one */
void ora_set_transfer_luts(const float *yuvxyb_eotf, const float *palette_eotf);
void ora_get_transfer_luts(float *yuvxyb_eotf, float *palette_eotf);
float ora_cbrtf(float x);                                                   /* yuvxyb-math cbrtf (FreeBSD msun) */
void ora_linear_rgb_to_xyb(const float rgb[3], float xyb[3]);               /* yuvxyb, before make_positive */
void ora_gaussian_coeffs(float n2[3], float d1[3], int *radius);            /* libjxl CreateRecursiveGaussian(1.5) */
void ora_blur_plane(const float *in, float *out, int w, int h);             /* H pass then V pass */

/* cogset Kmeans::new(points, k).clusters(): points[n][3] f64.  centres[k][3], assign[n].  Returns the
 * number of update iterations performed, or -1 where cogset would panic. */
int ora_kmeans(const double *points, int n, int k, double *centres, int *assign);

/* ssimulacra2::compute_frame_ssimulacra2 on two sRGB8 images (alpha ignored).  Optional outputs:
 * avg[6 scales][18] = per scale {ssim[6], edge[12]} plane averages. */
double ora_ssimulacra2_rgba8(const uint8_t *src_rgba, const uint8_t *dst_rgba, int w, int h, double *avg);
/* Positive-XYB planes of one image at every scale: out[87360*3], layout [scale][channel][y][x]. */
void ora_xyb_pyramid_rgba8(const uint8_t *rgba, int w, int h, float *out);
/* Source-side planes the GPU engine precomputes per image: mu1 = blur(i1), s11 = blur(i1*i1),
 * same [scale][channel][y][x] layout. */
void ora_source_planes_rgba8(const uint8_t *rgba, int w, int h, float *mu1, float *s11);

#ifdef __cplusplus
}
#endif
#endif
