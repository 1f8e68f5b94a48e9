/*
 * include/snesgpu.h -- C ABI of the B200-native palette-optimisation engine (libsnesgpu.so).
 *
 * The reference (aexoden/snesimage) has no FFI of its own: the hot path is the private
 * `OptimizedImage` method set in /root/reference/src/lib.rs.  This header is the seam a maintainer
 * would bind from Rust (`extern "C"` block): one entry point per
 * `OptimizedImage` method, plus batched entry points for the candidate loops inside
 * `optimize_palette_entry_{random,nes,channel}`.
 *
 * Conventions
 *   - every function returns 0 on success or a negative SNES_E_* code; snes_last_error() returns a
 *     thread-local message (the Rust side wraps it as anyhow!(msg).context(..), mirroring the
 *     `.context("Unable to optimize image")` strings of lib.rs:82,186,212,...).
 *   - host pointers are borrowed for the duration of the call only; the library owns all device
 *     memory.  Calls are synchronous unless the name ends in `_dev`.
 *   - a snes_ctx (and the images created from it) is not thread-safe, like the reference.
 *   - there is NO CPU fallback: every compute entry point fails with SNES_E_CUDA if no sm_100 GPU.
 */
#ifndef SNESGPU_H
#define SNESGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNES_OK 0
#define SNES_E_INVALID (-1)  /* bad argument (size, index, config) */
#define SNES_E_CUDA (-2)     /* CUDA runtime error / no device */
#define SNES_E_KMEANS (-3)   /* cogset's assert!(2 <= k && k < n) would panic */
#define SNES_E_NOMEM (-4)

#define SNES_WIDTH 256            /* lib.rs:29 */
#define SNES_HEIGHT 256           /* lib.rs:30 */
#define SNES_NES_COLOR_COUNT 56   /* lib.rs:31 */

typedef struct snes_ctx snes_ctx;
typedef struct snes_image snes_image; /* struct OptimizedImage, lib.rs:33-43 */

/* config::Config, /root/reference/src/config.rs:13-30 (the fields OptimizedImage::new consumes) */
typedef struct snes_config {
    int32_t subpalette_count;     /* -c, default 1 */
    int32_t subpalette_size;      /* -s, default 7 */
    uint8_t dither;               /* -d */
    uint8_t perceptual_palettes;  /* --perceptual-palettes */
    uint8_t nes;                  /* --nes */
    uint8_t reserved;
} snes_config;

/* (best error, candidate index) pair produced per image by the batched evaluators; 16 bytes so a
 * step's results for all images form one all-gather payload. */
typedef struct snes_best {
    double err;
    int32_t idx;  /* index into the candidate list, -1 if no candidate was evaluated */
    int32_t pad;
} snes_best;

/* One iteration of the optimiser loop of run() (lib.rs:889-933) names a palette entry (and, in channel mode, a channel). */
typedef struct snes_step {
    int32_t palette;   /* subpalette, 0 .. subpalette_count-1 */
    int32_t index;     /* entry within it, 0 .. subpalette_size-1 */
    int32_t channel;   /* 0 r, 1 g, 2 b: optimize_palette_entry_channel only */
    int32_t reserved;
} snes_step;

const char *snes_last_error(void);
int snes_version(void);

/* ---- context: one per GPU (one process per GPU) -------------------------------------------- */
int snes_ctx_create(int device, snes_ctx **out);
void snes_ctx_destroy(snes_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t snes_ctx_kernel_launches(const snes_ctx *ctx);
/* make the library enqueue on an external CUDA stream (e.g. torch's current stream); NULL = the context's own
 * non-blocking stream; pass cudaStreamLegacy ((void*)0x1) to name the legacy default stream */
int snes_ctx_set_stream(snes_ctx *ctx, void *cuda_stream);
int snes_ctx_synchronize(snes_ctx *ctx);
/* Per-launch device timing with CUDA events on the launching stream.  _begin clears and enables; _end
 * disables, synchronises and writes a JSON object {"<kernel>": {"ms": total, "n": launches}, ...}
 * (at most cap-1 bytes + NUL; *len = full length). */
int snes_ctx_profile_begin(snes_ctx *ctx);
/* restrict the event pairs to launches whose kernel name contains name_part (NULL or "": every launch).  Two event records
 * per launch are cheap but not free: a timed region of many short steps brackets its dominant kernel only. */
int snes_ctx_profile_only(snes_ctx *ctx, const char *name_part);
int snes_ctx_profile_end(snes_ctx *ctx, char *buf, size_t cap, size_t *len);
/* kernel selection (both scorers give the same f32 planes; kept switchable for A/B checks):
 *   fused             3 = k_score_v3 (default: persistent 4-warp CTAs, TMA staging, packed-f32 horizontal pass),
 *                     2 = k_score_v2 (its predecessor); anything else is refused
 *   block_width       ignored (the column block is 32)
 *   delta_assign != 0 without dithering, a candidate re-decides only the pixels its entry can change (default)
 * Env overrides at context creation: SNESGPU_FUSED, SNESGPU_DELTA. */
int snes_ctx_set_scorer(snes_ctx *ctx, int fused, int block_width, int delta_assign);
/* Replace the two 256-entry sRGB -> linear tables (index = 8-bit value): yuvxyb's transfer function, which feeds the
 * SSIMULACRA2 planes, and palette's Srgb::into_linear, which feeds Lab.  NULL keeps / restores the built-in table (libm
 * powf).  For pinning against tables dumped from the real crates (tests/golden/gen_reference_vectors.rs); call it before
 * any image is created -- images keep the planes they were built with.  The tables live in constant memory: they are shared
 * by every context of the process. */
int snes_ctx_set_transfer_luts(snes_ctx *ctx, const float *yuvxyb_eotf /* 256 or NULL */, const float *palette_eotf /* 256 or NULL */);
/* SSIMULACRA2's pooling table gives weight 0.0 to both ssim_map numbers of (channel X, scale 0) and of (channel B, scale 0);
 * k_score_v3 therefore computes only edge_diff_map there (one blurred plane per channel instead of three), decided from the table
 * itself at context creation.  on != 0: compute every term everywhere (A/B check: error() must not change by a bit). Env:
 * SNESGPU_ALL_TERMS=1. */
int snes_ctx_set_all_terms(snes_ctx *ctx, int on);
/* Self-check of the kernels' cube root (yuvxyb's cbrtf = FreeBSD msun s_cbrtf.c, the opsin transfer of linear_rgb_to_xyb): the
 * kernels evaluate it with one f32 and one division-free f64 Halley step and fall back to the restated msun function wherever the
 * f64 result is too close to an f32 rounding boundary for the two to be guaranteed equal, and for inputs outside [2^-126, 2^120)
 * (common.cuh: msun_cbrtf_fast).  This runs
 * both on every float whose bit pattern lies in [lo_bits, hi_bits) and counts the inputs on which they differ (must be 0) and the
 * inputs that took the fallback. */
int snes_ctx_cbrt_selfcheck(snes_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits, uint64_t *mismatches, uint64_t *fallbacks);
/* how many candidate evaluations have their scratch live at once (default 2048, env SNESGPU_CHUNK) */
int snes_ctx_set_chunk(snes_ctx *ctx, int evaluations);

/* ---- OptimizedImage ------------------------------------------------------------------------- */
/* OptimizedImage::new (lib.rs:46-65).  rgba: width*height*4 bytes, r,g,b,a.  Only 256x256 is accepted
 * (the reference's own tile table is fixed at 32x32, lib.rs:58,565).  Uploads the image and
 * precomputes the source-side SSIMULACRA2 planes. */
int snes_image_new(snes_ctx *ctx, const uint8_t *rgba, int width, int height, const snes_config *cfg, snes_image **out);
void snes_image_free(snes_image *im);

int snes_image_initialize_tiles(snes_image *im);     /* lib.rs:79-189 */
int snes_image_recalculate_palettes(snes_image *im); /* lib.rs:407-415 */
int snes_image_optimize(snes_image *im);             /* lib.rs:425-501 */
int snes_image_error(snes_image *im, double *err);   /* lib.rs:503-548 */
int snes_image_as_rgba(snes_image *im, uint8_t *out_rgba /* 65536*4 */); /* lib.rs:550-577 */
/* lib.rs:579-625 + serde_json's compact to_string(): writes at most cap bytes (NUL-terminated when it
 * fits) and always stores the required length (without NUL) in *len. */
int snes_image_as_json(snes_image *im, char *buf, size_t cap, size_t *len);

/* lib.rs:191-240; the 64 `rand::rng()` draws become the explicit list cand[ncand][3] (5-bit r,g,b) */
int snes_image_optimize_palette_entry_random(snes_image *im, int palette, int index, const uint8_t *cand, int ncand);
int snes_image_optimize_palette_entry_nes(snes_image *im, int palette, int index);                  /* lib.rs:242-284 */
int snes_image_optimize_palette_entry_channel(snes_image *im, int palette, int index, int channel); /* lib.rs:286-328 */

/* state accessors (fields of OptimizedImage / Palette) */
int snes_image_get_palette(snes_image *im, uint8_t *out /* C*S*3 */);
int snes_image_set_palette(snes_image *im, const uint8_t *in);
int snes_image_get_tile_palettes(snes_image *im, uint8_t *out /* 1024 */);
int snes_image_set_tile_palettes(snes_image *im, const uint8_t *in);
int snes_image_get_palette_map(snes_image *im, uint8_t *out /* 65536 */);
int snes_image_set_palette_map(snes_image *im, const uint8_t *in);
/* 64-bit FNV-1a over palette, tile_palettes and palette_map: replicas of an image on different ranks must agree on it */
int snes_image_state_checksum(snes_image *im, uint64_t *out);

/* ---- the same methods over a batch of images (independent OptimizedImages sharing one Config) -- */
int snes_batch_initialize_tiles(snes_ctx *ctx, snes_image *const *images, int nimg);     /* lib.rs:79-189 */
int snes_batch_recalculate_palettes(snes_ctx *ctx, snes_image *const *images, int nimg); /* lib.rs:407-415 */
int snes_batch_optimize(snes_ctx *ctx, snes_image *const *images, int nimg);             /* lib.rs:425-501 */
int snes_batch_error(snes_ctx *ctx, snes_image *const *images, int nimg, double *errors /* nimg, optional */);
/* asynchronous error(): refreshes each image's cached error on the context's stream; d_errors optional */
int snes_batch_error_dev(snes_ctx *ctx, snes_image *const *images, int nimg, double *d_errors);

/* ---- batched candidate evaluation: the loops at lib.rs:205-220, 252-262, 296-306 ------------- */
/* For every image j and candidate k: colours[palette*S+index] = cand[j][k]; optimize(); error().
 * The images' own state is not modified.  scores[nimg*ncand], maps[nimg*ncand*65536] and best[nimg]
 * are optional (NULL to skip).  best[j] is the strict-< first minimum over k (lib.rs:216). */
int snes_batch_eval_candidates(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                               const uint8_t *cand /* nimg*ncand*3 */, int ncand, double *scores, uint8_t *maps,
                               snes_best *best);
/* Same with device-resident candidate list / outputs, asynchronous on the context's stream.
 * cand_idx_base is added to every best[j].idx (a rank's offset into a sharded candidate list). */
int snes_batch_eval_candidates_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                   const uint8_t *d_cand, int ncand, int cand_idx_base, double *d_scores,
                                   snes_best *d_best);
/* The first two statements of optimize_palette_entry_{random,channel} in one pass: `best_error = self.error()`
 * (lib.rs:199, 294; refreshes every image's cached error, which snes_batch_apply_best_dev compares against) followed
 * by the candidate loop above.  The 3*nimg (image, channel) scoring items of error() share the candidates' scorer
 * launch instead of running as a launch of their own. */
int snes_batch_error_eval_candidates_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                         const uint8_t *d_cand, int ncand, int cand_idx_base, double *d_scores,
                                         snes_best *d_best);
/* The same for a rank of a candidate-sharded job: d_cand_all holds the full list, ncand_all candidates per image,
 * identical on every rank; this rank evaluates candidates [cand_lo, cand_lo + ncand) of every image in place (no copy of
 * the slice) and reports indices into the full list.  ncand == 0 (more ranks than candidates) is allowed: error() still
 * runs and the records carry idx = -1, which never wins snes_merge_best_dev. */
int snes_batch_error_eval_candidates_slice_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                               const uint8_t *d_cand_all, int ncand_all, int cand_lo, int ncand,
                                               double *d_scores, snes_best *d_best);
/* Accept rule of lib.rs:199,216-219,236-237 applied per image after the (possibly cross-rank) argmin:
 * if best[j].err < current error of image j, entry (palette,index) becomes cand_all[j][best[j].idx] and
 * the image is re-optimised; the image's cached error is updated.  d_cand_all holds ncand_all candidates
 * per image (the full, unsharded list).  Asynchronous on the context's stream. */
int snes_batch_apply_best_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                              const uint8_t *d_cand_all, int ncand_all, const snes_best *d_best);
/* Cross-rank argmin after an all-gather of every rank's best records: rank r's nimg records start at
 * d_gathered + r * rank_stride (rank_stride >= nimg; indices are already global); d_out[j] = lexicographic minimum of
 * (err, idx) over the nranks ranks, records with idx < 0 never winning (lib.rs:216 for any rank count). */
int snes_merge_best_dev(snes_ctx *ctx, const snes_best *d_gathered, int nranks, int rank_stride, int nimg, snes_best *d_out);
/* Host-buffer convenience for one whole optimiser step over many images (eval + argmin + accept). */
int snes_batch_step_random(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                           const uint8_t *cand /* nimg*ncand*3 */, int ncand, snes_best *best /* optional */,
                           double *errors_after /* nimg, optional */);

/* One optimize_palette_entry_random over a batch, sharded over the GPUs of one box, host buffers in and out (one process
 * per GPU).  The exchange in the middle -- an all-gather of every rank's d_best_local, 16 bytes per image
 * -- belongs to the caller's process group (NCCL); the library provides the two halves around it:
 *   _begin  copies cand (host, nimg * ncand_all * 3, identical on the ranks that share these images) to the device, runs
 *           error() and this rank's candidate slice [cand_lo, cand_lo + ncand), leaves nimg records in d_best_local
 *           (device memory of the caller, e.g. the send buffer of the all-gather).  Asynchronous on the context's stream.
 *   _end    d_gathered: the records of the nranks ranks that share these images, rank r at d_gathered + r * rank_stride;
 *           lexicographic merge (lib.rs:216 for any rank count), accept, optimize(); best[nimg] / errors_after[nimg]
 *           (host, optional) receive the merged records / error() of the new state; returns when the stream has drained.
 * Device-pointer candidate lists (the *_dev entry points) must hold components <= 32: larger values are clamped for the
 * evaluation, never applied, and reported by the next snes_ctx_synchronize() as SNES_E_INVALID. */
int snes_batch_step_random_shard_begin(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                       const uint8_t *cand, int ncand_all, int cand_lo, int ncand, snes_best *d_best_local);
int snes_batch_step_random_shard_end(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                     const snes_best *d_gathered, int nranks, int rank_stride, snes_best *best,
                                     double *errors_after);

/* ---- the same sharded step with the collective INSIDE the library: one call, host buffers in and out ---------------------
 * For hosts without a process-group library of their own (the reference is Rust).  NCCL is loaded at run time (libnccl.so.2).
 *   snes_comm_unique_id   on one rank: a 128-byte NCCL unique id, to be handed to every rank by any means
 *   snes_ctx_comm_init    on every rank (collective): the context's communicator over `world` ranks
 *   snes_dist_plan        which images of a job of nimg_total images a rank holds -- [img_lo, img_hi) -- and how its group slices
 *                         the candidates: as many image groups as divide the world and do not exceed nimg_total; the ranks of
 *                         a group (cand_ranks of them, this one being `slice`) hold replicas and split the candidate list;
 *                         slots = records per rank in the all-gather (the largest group)
 *   snes_dist_step_random optimize_palette_entry_random (lib.rs:191-240) for this rank's images as a member of the job: cand
 *                         (host, nimg_local * ncand * 3: the FULL lists of its images, identical on the ranks of its group) ->
 *                         error() + its slice of the candidates -> ncclAllGather of the records on the context's stream ->
 *                         merge over its group, accept, optimize().  best[nimg_local] (host, optional): the winning records;
 *                         all_best[world * slots] (host, optional): every rank's records as gathered.  Synchronous. */
int snes_comm_unique_id(uint8_t *out128);
int snes_ctx_comm_init(snes_ctx *ctx, const uint8_t *id128, int rank, int world);
int snes_ctx_comm_destroy(snes_ctx *ctx);
int snes_dist_plan(int nimg_total, int rank, int world, int *img_lo, int *img_hi, int *cand_ranks, int *slice, int *slots);
int snes_dist_step_random(snes_ctx *ctx, snes_image *const *images, int nimg_local, int nimg_total, int palette, int index,
                          const uint8_t *cand, int ncand, snes_best *best, snes_best *all_best);

/* optimize_palette_entry_nes / _channel for every image of a batch (lib.rs:242-284, 286-328) */
int snes_batch_step_nes(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index, snes_best *best,
                        double *errors_after);
int snes_batch_step_channel(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index, int channel,
                            snes_best *best, double *errors_after);

/* ---- several palette entries per launch (lib.rs:889-933 looked at as a whole; TODO.md:33-35) --------- */
/* Candidates of nsteps palette entries evaluated against ONE state in one launch sequence: for image j, step s and candidate
 * k: colours[steps[s]] = cand[j][s][k]; optimize(); error().  cand: nimg*nsteps*ncand*3; scores[nimg*nsteps*ncand] and
 * best[nimg*nsteps] (strict-< first minimum of each step's list, idx within that list) are optional.  The images' own state
 * is not modified.  With every entry of the palette as a step this is the whole-sweep batch TODO.md:33-35 wishes for. */
int snes_batch_eval_candidates_multi(snes_ctx *ctx, snes_image *const *images, int nimg, const snes_step *steps, int nsteps,
                                     const uint8_t *cand, int ncand, double *scores, snes_best *best);
/* nsteps consecutive iterations of the loop body of run() (lib.rs:889-910: optimize_palette_entry_* + optimize() + error())
 * in one call, exactly the reference's trajectory: all steps' candidates are evaluated against the current state, then the
 * steps are taken in order up to and including the first one that accepts a candidate (an iteration that accepts nothing
 * leaves the state it was evaluated against, so evaluating the next one early changes nothing).  mode 0: random, cand =
 * nsteps*ncand*3 explicit colours (the reference's rand::rng() draws, lib.rs:205-208); 1: NES (a step always takes its first
 * minimum, lib.rs:250, and ends the run only when that changes the entry's colour); 2: channel (32 values of steps[s].channel).  *consumed = iterations the call stands for
 * (the caller advances its cursor by that many and drops the rest of the list), *error_before = error() of the state the
 * call started from (what every iteration before the accepting one ends with; in NES mode NaN unless such an iteration
 * exists), *error_after = error() of the new state. */
int snes_image_iterate(snes_image *im, int mode, const snes_step *steps, int nsteps, const uint8_t *cand, int ncand,
                       int *consumed, double *error_before, double *error_after);

/* ---- tile reassignment as evaluated candidates ------------------------------------------------ */
/* The reference changes a tile's subpalette only by hand (a click cycles it, lib.rs:1005-1017; TODO.md:36-37 wishes for
 * an automatic version).  Here a move is a candidate like a palette colour: for image j and move k,
 * tile_palettes[moves[j][k][0]] = moves[j][k][1]; optimize(); error().  moves: nimg*nmoves pairs (tile index =
 * tile_y*32 + tile_x, subpalette).  scores[nimg*nmoves], maps[nimg*nmoves*65536], best[nimg] optional; the images' own
 * state is not modified. */
int snes_batch_eval_tile_moves(snes_ctx *ctx, snes_image *const *images, int nimg, const int32_t *moves, int nmoves,
                               double *scores, uint8_t *maps, snes_best *best);
/* Evaluate, then apply each image's best move if it is strictly better than the image's error() (the accept rule of
 * lib.rs:216-219) and optimize().  applied[nimg] (optional): 1 where a move was taken. */
int snes_batch_step_tile_moves(snes_ctx *ctx, snes_image *const *images, int nimg, const int32_t *moves, int nmoves,
                               snes_best *best, uint8_t *applied);

/* ---- colour primitives exposed for tests (lib.rs:628-795, 1080-1100) ------------------------- */
/* Nearest entry for n targets (f64 triples) against a list of ncolors 5-bit colours on the GPU. */
int snes_closest_color_index(snes_ctx *ctx, const uint8_t *colors5, int ncolors, const double *targets, int n,
                             int cielab, int32_t *out_index);
/* SnesColor::new_nes_only (lib.rs:640-660) for n colours. */
int snes_new_nes_only(snes_ctx *ctx, const uint8_t *colors5, int n, int cielab, uint8_t *out5);
/* Debug/test taps: positive-XYB pyramid and blurred source planes of an image, [scale][ch][y][x] f32,
 * 87360*3 floats each. */
int snes_image_debug_planes(snes_image *im, float *xyb, float *mu1, float *s11);
/* per-pixel Lab<D65,f32> of the original (perceptual_palettes only), 65536*4 floats (L,a,b,0) */
int snes_image_debug_lab(snes_image *im, float *lab);
/* centres[256*3] and status[512] (status[p] per problem, status[C+p] = Lloyd iterations) of the last k-means */
int snes_image_kmeans_debug(snes_image *im, double *centres, int32_t *status);

#ifdef __cplusplus
}
#endif
#endif
