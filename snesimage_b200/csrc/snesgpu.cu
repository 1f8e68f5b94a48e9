// libsnesgpu.so -- C ABI (include/snesgpu.h) over the sm_100a kernels.
//
// Host side only orchestrates: it owns device buffers, builds the small constant tables (sRGB
// transfer LUTs, recursive-Gaussian taps, pooling weights, NES colours), launches kernels on one
// stream and copies state in and out.  Every arithmetic step of the hot path runs on the GPU; there
// is no CPU fallback -- without a usable sm_100 device every compute entry point returns SNES_E_CUDA.
#include "../../include/snesgpu.h"

#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dither.cuh"
#include "kernels.cuh"
#include "kmeans.cuh"
#include "lab.cuh"
#include "score_v2.cuh"
#include "score_v3.cuh"
#include "assign_delta.cuh"

using namespace snes;

static_assert(sizeof(Best) == sizeof(snes_best), "snes_best layout");

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(SNES_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)
#define RET(call)                 \
    do {                          \
        int r_ = (call);          \
        if (r_ != SNES_OK) return r_; \
    } while (0)

extern "C" const char *snes_last_error(void) { return g_err.c_str(); }
extern "C" int snes_version(void) { return 1; }

// ------------------------------------------------------------------------------------------------
// context / image
// ------------------------------------------------------------------------------------------------
struct snes_ctx {
    int device = 0;
    cudaStream_t own = nullptr, stream = nullptr;
    cudaStream_t side = nullptr;              // k_score_v3 runs here next to k_score_pair on `stream` (fork / join by events)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int64_t launches = 0;
    int chunk = 2048;  // evaluations whose scratch (palette_map, coarse XYB pyramid) is live at once
    int fused = 3;    // 3: k_score_v3 (persistent 4-warp CTAs); 2: k_score_v2, its predecessor, kept as the A/B check (SNESGPU_FUSED)
    int nsm = 148;
    int *v3_counter = nullptr;        // [2]: work counters of k_score_v3 and k_score_pair
    float *v3_scratch = nullptr;
    float *self_xyb = nullptr;        // [img_cap][EVAL_XYB_FLOATS] coarse pyramid of the images' own state (fused error + candidates)
    double *self_partials = nullptr;  // [img_cap][NSCALES*3*NSUMS]
    bool no_cluster_kmeans = false;   // env SNESGPU_NO_CLUSTER_KMEANS=1: one CTA per k-means problem whatever their number (A/B runs)
    int pair_xb = 0;  // scale 0 of channels X and B carries no ssim_map weight in the pooling table: edge-only pair items (score_v3.cuh)
    int delta = 1;    // 1: no-dither candidates re-decide only the pixels the replaced entry can change (SNESGPU_DELTA)

    // per-chunk scratch
    size_t chunk_cap = 0;
    float *xyb_rm = nullptr;
    float *hbuf = nullptr;   // H planes of one image's source-side blur (image creation); also the as_rgba staging buffer
    int *d_fault = nullptr;  // set by k_tables when a device-resident candidate list holds a colour component > 32
    uint8_t *maps = nullptr;
    // per-batch scratch
    size_t eval_cap = 0;
    double *partials = nullptr, *scores = nullptr;
    CandEntry *cents = nullptr;
    uint8_t *cand = nullptr;
    size_t img_cap = 0;
    ImgDev *d_imgs = nullptr;
    ImgTm *d_imgtm = nullptr;   // tensor maps of the bound images, parallel to d_imgs
    KmScratch *d_km = nullptr;
    Best *best = nullptr;
    double *self_scores = nullptr;
    std::vector<snes_image *> cached;

    int *d_ints = nullptr;     // small integer scratch of the multi-entry calls: slots, channels, consumed, chosen
    size_t ints_cap = 0;
    Best *best_m = nullptr;    // [nimg * nsteps] first minima of a multi-entry call
    size_t best_m_cap = 0;
    bool maps_live = false;    // the scratch maps hold the gi maps of every evaluation of the last candidate step (one entry,
    int maps_nimg = 0, maps_ncand = 0, maps_slot = -1;   // full candidate lists): [maps_nimg][maps_ncand][NPIX]
    const uint8_t *maps_cand = nullptr;
    std::vector<snes_image *> maps_images;
    void *nccl_comm = nullptr; // ncclComm_t of snes_ctx_comm_init (the library's own communicator, NCCL loaded at run time)
    int comm_rank = 0, comm_world = 1;
    Best *d_send = nullptr, *d_gather = nullptr;   // [slots], [world * slots] records of snes_dist_step_random
    size_t dist_cap = 0;
    int shard_ncand_all = 0;   // candidates per image of the list snes_batch_step_random_shard_begin left in `cand`

    float4 *labtab = nullptr;  // BGR555 -> Lab<D65,f32>

    // per-launch CUDA-event timing (snes_ctx_profile_begin/end)
    bool profiling = false, prof_skip = false;
    std::string prof_filter;   // non-empty: only launches whose name contains it are bracketed by events
    std::vector<cudaEvent_t> prof_events;
    std::vector<const char *> prof_names;
    size_t prof_used = 0;
};

struct snes_image {
    snes_ctx *ctx = nullptr;
    snes_config cfg{};
    ImgDev dev{};
    KmScratch km{};
    void *slab = nullptr, *km_slab = nullptr;
    ImgTm tm{};                  // TMA tensor maps over the slab (source pyramid per scale, own palette_map)
    std::vector<uint8_t> alpha;  // host copy of the alpha channel (as_json)
    bool map_fresh = false;      // palette_map is optimize() of the current palette and tile assignment (no setter since)
};

// ------------------------------------------------------------------------------------------------
// TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*tm_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tm_encode_fn tm_encoder() {
    static tm_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (tm_encode_fn)p;
    }
    return fn;
}
// rank-`rank` tensor of f32 (esize 4) or u32 (esize 0) elements at `base`: dims[] elements, strides[] bytes for dims 1.., box[] elements
static int tm_make(CUtensorMap *out, const void *base, int esize, int rank, const uint64_t *dims, const uint64_t *strides, const uint32_t *box) {
    tm_encode_fn enc = tm_encoder();
    if (!enc) return fail(SNES_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gs[i - 1] = strides[i - 1];
    }
    const CUresult r = enc(out, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, (cuuint32_t)rank, const_cast<void *>(base), gd,
                           gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SNES_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return SNES_OK;
}
static uint32_t tm_box_rows(int s) { return (uint32_t)(((W >> s) < V3Smem::HB ? (W >> s) : V3Smem::HB) + 4); }
// an image's own planes: source pyramid scale s as (D, D, 3) f32, palette_map as (256, 256) u8
static int tm_make_image(ImgTm *tm, const ImgDev &dev) {
    for (int s = 0; s < NSCALES; s++) {
        const uint64_t D = W >> s, dims[3] = {D, D, 3}, strides[2] = {D * 4, D * D * 4};
        const uint32_t box[3] = {(uint32_t)V3Smem::IP, tm_box_rows(s), 1};
        RET(tm_make(&tm->src[s], dev.xyb_rm + 3 * (size_t)scale_off(s), 4, 3, dims, strides, box));
    }
    // palette_map rows as 4-pixel words (the conversion reads words; x coordinates are multiples of 4)
    const uint64_t dims[2] = {W / 4, H}, strides[1] = {W};
    const uint32_t box[2] = {V3Smem::RAWP, tm_box_rows(0)};
    return tm_make(&tm->own, dev.map, 0, 2, dims, strides, box);
}
// an evaluation buffer: palette_maps (may be null) as (256, 256, E) u8, coarse pyramids per scale as (D, D, 3, E) f32
static int tm_make_evals(EvalTm *tm, const uint8_t *maps, const float *xyb, int E) {
    memset(tm, 0, sizeof(*tm));
    if (maps) {
        const uint64_t dims[3] = {W / 4, H, (uint64_t)E}, strides[2] = {W, (uint64_t)NPIX};
        const uint32_t box[3] = {V3Smem::RAWP, tm_box_rows(0), 1};
        RET(tm_make(&tm->t[0], maps, 0, 3, dims, strides, box));
    }
    for (int s = 1; s < NSCALES; s++) {
        const uint64_t D = W >> s, dims[4] = {D, D, 3, (uint64_t)E}, strides[3] = {D * 4, D * D * 4, (uint64_t)EVAL_XYB_FLOATS * 4};
        const uint32_t box[4] = {(uint32_t)V3Smem::IP, tm_box_rows(s), 1, 1};
        RET(tm_make(&tm->t[s], xyb + 3 * (size_t)scale_off(s), 4, 4, dims, strides, box));
    }
    return SNES_OK;
}

static void prof_begin(snes_ctx *ctx, const char *name) {
    if (!ctx->profiling) return;
    ctx->prof_skip = !ctx->prof_filter.empty() && !strstr(name, ctx->prof_filter.c_str());
    if (ctx->prof_skip) return;
    cudaStream_t st = ctx->stream;
    if (ctx->prof_used + 2 > ctx->prof_events.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        ctx->prof_events.push_back(a);
        ctx->prof_events.push_back(b);
    }
    ctx->prof_names.push_back(name);
    cudaEventRecord(ctx->prof_events[ctx->prof_used], st);
}
static void prof_end(snes_ctx *ctx) {
    if (!ctx->profiling || ctx->prof_skip) return;
    cudaEventRecord(ctx->prof_events[ctx->prof_used + 1], ctx->stream);
    ctx->prof_used += 2;
}

// Every kernel launch goes through LAUNCH: it counts the launch and, while profiling is on, brackets it
// with CUDA events on the launching stream (per-kernel device time for bench.py's roofline object).
#define LAUNCH(ctx, name, ...)          \
    do {                                \
        prof_begin((ctx), (name));      \
        __VA_ARGS__;                    \
        prof_end((ctx));                \
        (ctx)->launches++;              \
        CK(cudaGetLastError());         \
    } while (0)

static int set_device(snes_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    return SNES_OK;
}

template <typename T>
static int dev_alloc(T **p, size_t n) {
    CK(cudaMalloc((void **)p, n * sizeof(T)));
    return SNES_OK;
}

static float srgb_eotf_yuvxyb(float x) {  // yuvxyb transfer: sRGB -> linear
    const float alpha = 1.0550107f, beta = 0.0030412825f;
    x = x > 0.0f ? x : 0.0f;
    if (x < 12.92f * beta) return x / 12.92f;
    return powf((x + (alpha - 1.0f)) / alpha, 2.4f);
}

static void inv3x3(double m[9]) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double inv = 1.0 / (a * A + b * B + c * C);
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

// libjxl CreateRecursiveGaussian(sigma = 1.5): Charalampidis 2016, three second-order sections.
static void gaussian_taps(float n2[3], float d1[3]) {
    const double sigma = 1.5, pi = 3.141592653589793238;
    const double radius = std::round(3.2795 * sigma + 0.2546);
    const double w = pi / (2.0 * radius);
    const double omega[3] = {w, 3.0 * w, 5.0 * w};
    const double p1 = 1.0 / std::tan(0.5 * omega[0]), p3 = -1.0 / std::tan(0.5 * omega[1]), p5 = 1.0 / std::tan(0.5 * omega[2]);
    const double r1 = p1 * p1 / std::sin(omega[0]), r3 = -p3 * p3 / std::sin(omega[1]), r5 = p5 * p5 / std::sin(omega[2]);
    double rho[3];
    for (int i = 0; i < 3; i++) rho[i] = std::exp(-0.5 * sigma * sigma * omega[i] * omega[i]) * (1.0 / radius);
    const double D13 = p1 * r3 - r1 * p3, D35 = p3 * r5 - r3 * p5, D51 = p5 * r1 - r5 * p1;
    const double rd13 = 1.0 / D13, z15 = D35 * rd13, z35 = D51 * rd13;
    double A[9] = {p1, p3, p5, r1, r3, r5, z15, z35, 1.0};
    inv3x3(A);
    const double gamma[3] = {1.0, radius * radius - sigma * sigma, z15 * rho[0] + z35 * rho[1] + rho[2]};
    for (int i = 0; i < 3; i++) {
        const double beta = A[3 * i] * gamma[0] + A[3 * i + 1] * gamma[1] + A[3 * i + 2] * gamma[2];
        n2[i] = (float)(-beta * std::cos(omega[i] * (radius + 1.0)));
        d1[i] = (float)(-2.0 * std::cos(omega[i]));
    }
}

static const double kWeights[108] = {
    0.0, 0.0007376606707406586, 0.0, 0.0, 0.0007793481682867309, 0.0, 0.0, 0.0004371155730107379, 0.0,
    1.1041726426657346, 0.00066284834129271, 0.00015231632783718752, 0.0, 0.0016406437456599754, 0.0,
    1.8422455520539298, 11.441172603757666, 0.0, 0.0007989109436015163, 0.000176816438078653, 0.0,
    1.8787594979546387, 10.94906990605142, 0.0, 0.0007289346991508072, 0.9677937080626833, 0.0,
    0.00014003424285435884, 0.9981766977854967, 0.00031949755934435053, 0.0004550992113792063, 0.0, 0.0,
    0.0013648766163243398, 0.0, 0.0, 0.0, 0.0, 0.0, 7.466890328078848, 0.0, 17.445833984131262,
    0.0006235601634041466, 0.0, 0.0, 6.683678146179332, 0.00037724407979611296, 1.027889937768264,
    225.20515300849274, 0.0, 0.0, 19.213238186143016, 0.0011401524586618361, 0.001237755635509985,
    176.39317598450694, 0.0, 0.0, 24.43300999870476, 0.28520802612117757, 0.0004485436923833408, 0.0, 0.0, 0.0,
    34.77906344483772, 44.835625328877896, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0008680556573291698, 0.0, 0.0, 0.0,
    0.0, 0.0, 0.0005313191874358747, 0.0, 0.00016533814161379112, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0004179171803251336,
    0.0017290828234722833, 0.0, 0.0020827005846636437, 0.0, 0.0, 8.826982764996862, 23.19243343998926, 0.0,
    95.1080498811086, 0.9863978034400682, 0.9834382792465353, 0.0012286405048278493, 171.2667255897307,
    0.9807858872435379, 0.0, 0.0, 0.0, 0.0005130064588990679, 0.0, 0.00010854057858411537};

static const uint8_t kNes[NES_COUNT][3] = {  // lib.rs:687-742
    {13, 13, 13}, {0, 2, 16},   {3, 0, 17},   {7, 0, 15},   {10, 0, 10},  {11, 0, 3},   {9, 2, 0},    {7, 3, 0},
    {4, 6, 0},    {0, 7, 0},    {0, 8, 0},    {0, 7, 4},    {0, 5, 10},   {0, 0, 0},    {23, 23, 23}, {3, 10, 24},
    {9, 6, 28},   {14, 4, 26},  {18, 3, 21},  {19, 5, 11},  {19, 6, 0},   {15, 9, 0},   {11, 12, 0},  {4, 14, 0},
    {0, 15, 0},   {0, 14, 8},   {0, 13, 17},  {0, 0, 0},    {31, 31, 31}, {13, 20, 31}, {17, 19, 31}, {22, 16, 31},
    {27, 14, 31}, {28, 14, 23}, {28, 17, 13}, {26, 19, 5},  {22, 21, 1},  {15, 24, 2},  {10, 25, 8},  {8, 25, 16},
    {8, 24, 24},  {9, 9, 9},    {31, 31, 31}, {25, 29, 31}, {27, 27, 31}, {29, 27, 31}, {31, 26, 31}, {31, 26, 30},
    {31, 27, 25}, {31, 28, 22}, {30, 30, 21}, {27, 31, 21}, {25, 31, 23}, {24, 31, 26}, {24, 30, 30}, {23, 24, 23}};

static int weights_allow_pair_xb() {
    return kWeights[(0 * 6 + 0) * 6 + 0] == 0.0 && kWeights[(0 * 6 + 0) * 6 + 3] == 0.0 && kWeights[(2 * 6 + 0) * 6 + 0] == 0.0 &&
           kWeights[(2 * 6 + 0) * 6 + 3] == 0.0;
}

static constexpr int kBlurHSmem = 4 * 2 * 32 * 33 * (int)sizeof(float);

static int ctx_init(snes_ctx *ctx, int nsm);

extern "C" int snes_ctx_create(int device, snes_ctx **out) {
    if (!out) return fail(SNES_E_INVALID, "snes_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SNES_E_CUDA, "snes_ctx_create: no such CUDA device");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SNES_E_CUDA, std::string("snes_ctx_create: libsnesgpu is built for sm_100a only; device is ") + prop.name);
    snes_ctx *ctx = new snes_ctx();
    ctx->device = device;
    const int rc = ctx_init(ctx, prop.multiProcessorCount);
    if (rc != SNES_OK) {   // g_err holds the message of the step that failed
        const std::string msg = g_err;
        snes_ctx_destroy(ctx);
        return fail(rc, msg);
    }
    *out = ctx;
    return SNES_OK;
}

static int ctx_init(snes_ctx *ctx, int nsm) {
    const int device = ctx->device;
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking));
    ctx->stream = ctx->own;
    CK(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    if (const char *c = getenv("SNESGPU_CHUNK")) {
        const int v = atoi(c);
        if (v > 0) ctx->chunk = v;
    }

    if (const char *c = getenv("SNESGPU_FUSED")) ctx->fused = atoi(c) == 2 ? 2 : 3;
    if (const char *c = getenv("SNESGPU_DELTA")) ctx->delta = atoi(c) != 0;
    if (const char *c = getenv("SNESGPU_NO_CLUSTER_KMEANS")) ctx->no_cluster_kmeans = atoi(c) != 0;

    float lut[256], lut2[256], n2[3], d1[3];
    for (int v = 0; v < 256; v++) {
        lut[v] = srgb_eotf_yuvxyb((float)v / 255.0f);
        const float c = (float)v / 255.0f;  // palette: Srgb<u8>::into_format, Srgb::into_linear
        lut2[v] = c <= 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
    }
    gaussian_taps(n2, d1);
    uint8_t nes4[NES_COUNT][4];
    for (int i = 0; i < NES_COUNT; i++) {
        nes4[i][0] = kNes[i][0];
        nes4[i][1] = kNes[i][1];
        nes4[i][2] = kNes[i][2];
        nes4[i][3] = 0;
    }
    CK(cudaMemcpyToSymbol(c_lin_lut, lut, sizeof(lut)));
    CK(cudaMemcpyToSymbol(c_srgb_lin_lut, lut2, sizeof(lut2)));
    CK(cudaMemcpyToSymbol(c_n2, n2, sizeof(n2)));
    CK(cudaMemcpyToSymbol(c_d1, d1, sizeof(d1)));
    {
        float2 n2p[3], d1p[3], md1p[3];
        float md1[3];
        for (int k = 0; k < 3; k++) {
            n2p[k] = make_float2(n2[k], n2[k]);
            d1p[k] = make_float2(d1[k], d1[k]);
            md1p[k] = make_float2(-d1[k], -d1[k]);
            md1[k] = -d1[k];
        }
        CK(cudaMemcpyToSymbol(c_n2p, n2p, sizeof(n2p)));
        CK(cudaMemcpyToSymbol(c_d1p, d1p, sizeof(d1p)));
        CK(cudaMemcpyToSymbol(c_md1p, md1p, sizeof(md1p)));
        CK(cudaMemcpyToSymbol(c_md1, md1, sizeof(md1)));
        const float2 m1p = make_float2(-1.0f, -1.0f);
        CK(cudaMemcpyToSymbol(c_m1p, &m1p, sizeof(m1p)));
    }
    CK(cudaMemcpyToSymbol(c_weight, kWeights, sizeof(kWeights)));
    // weights of (channel c, scale s): kWeights[(c * 6 + s) * 6 + {0: ssim mean, 1: artifact mean, 2: detail mean, 3: ssim 4-norm, ...}].
    // ssim_map of a (channel, scale) whose two ssim weights are exactly zero cannot change the score; the scorer then skips the two
    // blur planes only ssim_map needs.  Decided from the table itself, so a corrected table changes the decision with it.
    ctx->pair_xb = weights_allow_pair_xb();
    if (const char *c = getenv("SNESGPU_ALL_TERMS"))
        if (atoi(c) != 0) ctx->pair_xb = 0;
    CK(cudaMemcpyToSymbol(c_nes, nes4, sizeof(nes4)));
    CK(cudaFuncSetAttribute(k_blur_h, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlurHSmem));
    CK(cudaFuncSetAttribute(k_score_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(V2Smem)));
    CK(cudaFuncSetAttribute(k_score_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(V3Smem)));
    CK(cudaFuncSetAttribute(k_score_v3, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ctx->nsm = nsm;
    RET(dev_alloc(&ctx->v3_counter, 2));
    CK(cudaFuncSetAttribute(k_score_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(V3PairSmem)));
    CK(cudaFuncSetAttribute(k_score_pair, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    RET(dev_alloc(&ctx->hbuf, (size_t)EVAL_XYB_FLOATS * 2));
    RET(dev_alloc(&ctx->d_fault, 1));
    CK(cudaMemsetAsync(ctx->d_fault, 0, sizeof(int), ctx->stream));
    RET(dev_alloc(&ctx->v3_scratch, (size_t)ctx->nsm * (V3_CTAS_PER_SM + V3_PAIR_CTAS_PER_SM) * V3_HSCRATCH_FLOATS));

    RET(dev_alloc(&ctx->labtab, 32768));
    LAUNCH(ctx, "k_build_lab_table", k_build_lab_table<<<128, 256, 0, ctx->stream>>>(ctx->labtab));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNES_OK;
}

static void free_scratch(snes_ctx *ctx) {
    cudaFree(ctx->xyb_rm);
    cudaFree(ctx->maps);
    cudaFree(ctx->partials);
    cudaFree(ctx->scores);
    cudaFree(ctx->cents);
    cudaFree(ctx->cand);
    cudaFree(ctx->d_imgs);
    cudaFree(ctx->d_imgtm);
    ctx->d_imgtm = nullptr;
    cudaFree(ctx->d_km);
    cudaFree(ctx->best);
    cudaFree(ctx->self_scores);
    cudaFree(ctx->self_xyb);
    cudaFree(ctx->self_partials);
}

extern "C" void snes_ctx_destroy(snes_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_scratch(ctx);
    snes_ctx_comm_destroy(ctx);
    cudaFree(ctx->d_send);
    cudaFree(ctx->d_gather);
    cudaFree(ctx->labtab);
    cudaFree(ctx->d_ints);
    cudaFree(ctx->best_m);
    cudaFree(ctx->hbuf);
    cudaFree(ctx->d_fault);
    cudaFree(ctx->v3_counter);
    cudaFree(ctx->v3_scratch);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->own) cudaStreamDestroy(ctx->own);
    delete ctx;
}

extern "C" int64_t snes_ctx_kernel_launches(const snes_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int snes_ctx_set_stream(snes_ctx *ctx, void *cuda_stream) {
    if (!ctx) return fail(SNES_E_INVALID, "ctx is NULL");
    RET(set_device(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own;
    return SNES_OK;
}

extern "C" int snes_ctx_synchronize(snes_ctx *ctx) {
    if (!ctx) return fail(SNES_E_INVALID, "ctx is NULL");
    RET(set_device(ctx));
    int fault = 0;
    CK(cudaMemcpyAsync(&fault, ctx->d_fault, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (fault) {
        CK(cudaMemsetAsync(ctx->d_fault, 0, sizeof(int), ctx->stream));
        return fail(SNES_E_INVALID, "a device-resident candidate list held a colour component > 32 (clamped for evaluation, never applied)");
    }
    return SNES_OK;
}

// The two 256-entry sRGB -> linear tables are the only place where the third-party transfer functions enter the GPU path
// (yuvxyb's for the SSIMULACRA2 planes, palette's for Lab).  The built-in ones come from libm's powf; a maintainer who can
// build the crates dumps theirs (tests/golden/gen_reference_vectors.rs) and swaps them in here -- no kernel changes.
extern "C" int snes_ctx_set_transfer_luts(snes_ctx *ctx, const float *yuvxyb_eotf, const float *palette_eotf) {
    if (!ctx) return fail(SNES_E_INVALID, "ctx is NULL");
    RET(set_device(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    float lut[256], lut2[256];
    for (int v = 0; v < 256; v++) {
        lut[v] = yuvxyb_eotf ? yuvxyb_eotf[v] : srgb_eotf_yuvxyb((float)v / 255.0f);
        const float c = (float)v / 255.0f;
        lut2[v] = palette_eotf ? palette_eotf[v] : (c <= 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f));
    }
    CK(cudaMemcpyToSymbol(c_lin_lut, lut, sizeof(lut)));
    CK(cudaMemcpyToSymbol(c_srgb_lin_lut, lut2, sizeof(lut2)));
    LAUNCH(ctx, "k_build_lab_table", k_build_lab_table<<<128, 256, 0, ctx->stream>>>(ctx->labtab));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNES_OK;
}

extern "C" int snes_ctx_profile_only(snes_ctx *ctx, const char *name_part) {
    if (!ctx) return fail(SNES_E_INVALID, "ctx is NULL");
    ctx->prof_filter = name_part ? name_part : "";
    return SNES_OK;
}

extern "C" int snes_ctx_profile_begin(snes_ctx *ctx) {
    if (!ctx) return fail(SNES_E_INVALID, "ctx is NULL");
    ctx->profiling = true;
    ctx->prof_names.clear();
    ctx->prof_used = 0;
    return SNES_OK;
}

// Stops profiling, drains the stream and writes {"kernel": {"ms": total, "n": launches}, ...} as JSON.
extern "C" int snes_ctx_profile_end(snes_ctx *ctx, char *buf, size_t cap, size_t *len) {
    if (!ctx || !len) return fail(SNES_E_INVALID, "snes_ctx_profile_end: NULL argument");
    RET(set_device(ctx));
    ctx->profiling = false;
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<std::string> names;
    std::vector<double> ms;
    std::vector<long> cnt;
    for (size_t i = 0; i < ctx->prof_names.size(); i++) {
        float t = 0.0f;
        CK(cudaEventElapsedTime(&t, ctx->prof_events[2 * i], ctx->prof_events[2 * i + 1]));
        size_t k = 0;
        while (k < names.size() && names[k] != ctx->prof_names[i]) k++;
        if (k == names.size()) {
            names.push_back(ctx->prof_names[i]);
            ms.push_back(0.0);
            cnt.push_back(0);
        }
        ms[k] += t;
        cnt[k]++;
    }
    std::string s = "{";
    char tmp[128];
    for (size_t k = 0; k < names.size(); k++) {
        snprintf(tmp, sizeof tmp, "%s\"%s\": {\"ms\": %.6f, \"n\": %ld}", k ? ", " : "", names[k].c_str(), ms[k], cnt[k]);
        s += tmp;
    }
    s += "}";
    ctx->prof_names.clear();
    ctx->prof_used = 0;
    *len = s.size();
    if (buf && cap > 0) {
        const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(buf, s.data(), n);
        buf[n] = '\0';
    }
    return SNES_OK;
}

extern "C" int snes_ctx_set_scorer(snes_ctx *ctx, int fused, int block_width, int delta_assign) {
    (void)block_width;   // the column block is fixed at 32 since k_score_fused<16/32> left the build
    if (!ctx || (fused != 2 && fused != 3)) return fail(SNES_E_INVALID, "snes_ctx_set_scorer: scorer must be 3 (k_score_v3) or 2 (k_score_v2)");
    ctx->fused = fused;
    ctx->delta = delta_assign != 0;
    return SNES_OK;
}

extern "C" int snes_ctx_set_all_terms(snes_ctx *ctx, int on) {
    if (!ctx) return fail(SNES_E_INVALID, "ctx is NULL");
    ctx->pair_xb = on ? 0 : weights_allow_pair_xb();
    return SNES_OK;
}

// every float with bit pattern in [lo_bits, hi_bits): msun_cbrtf against msun_cbrtf_fast
__global__ void k_cbrt_selfcheck(uint32_t lo_bits, uint32_t hi_bits, unsigned long long *mismatches, unsigned *fallbacks) {
    unsigned long long bad = 0;
    for (unsigned long long b = lo_bits + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < hi_bits; b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)b);
        bad += __float_as_uint(msun_cbrtf(x)) != __float_as_uint(msun_cbrtf_fast(x, fallbacks));
    }
    if (bad) atomicAdd(mismatches, bad);
}

extern "C" int snes_ctx_cbrt_selfcheck(snes_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits, uint64_t *mismatches, uint64_t *fallbacks) {
    if (!ctx || !mismatches || !fallbacks || hi_bits < lo_bits) return fail(SNES_E_INVALID, "snes_ctx_cbrt_selfcheck: bad argument");
    unsigned long long *d = nullptr;
    CK(cudaMalloc(&d, 16));
    CK(cudaMemsetAsync(d, 0, 16, ctx->stream));
    k_cbrt_selfcheck<<<148 * 8, 256, 0, ctx->stream>>>(lo_bits, hi_bits, d, reinterpret_cast<unsigned *>(d + 1));
    unsigned long long h[2] = {0, 0};
    const cudaError_t e = cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CK(e);
    CK(e2);
    *mismatches = h[0];
    *fallbacks = h[1] & 0xffffffffull;
    return SNES_OK;
}

extern "C" int snes_ctx_set_chunk(snes_ctx *ctx, int evaluations) {
    if (!ctx || evaluations < 1) return fail(SNES_E_INVALID, "snes_ctx_set_chunk: bad argument");
    ctx->chunk = evaluations;
    return SNES_OK;
}

// ---- scratch management ------------------------------------------------------------------------
// Scratch of one chunk of evaluations: the palette_map (64 KiB) and the coarse XYB pyramid per evaluation.
static int ensure_chunk(snes_ctx *ctx, size_t n) {
    if (n > ctx->chunk_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->xyb_rm);
        cudaFree(ctx->maps);
        ctx->xyb_rm = nullptr;
        ctx->maps = nullptr;
        ctx->chunk_cap = 0;
        RET(dev_alloc(&ctx->xyb_rm, n * EVAL_XYB_FLOATS));
        RET(dev_alloc(&ctx->maps, n * NPIX));
        ctx->chunk_cap = n;
    }
    return SNES_OK;
}

static int ensure_evals(snes_ctx *ctx, size_t n) {
    if (n <= ctx->eval_cap) return SNES_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->partials);
    cudaFree(ctx->scores);
    cudaFree(ctx->cents);
    cudaFree(ctx->cand);
    ctx->partials = ctx->scores = nullptr;
    ctx->cents = nullptr;
    ctx->cand = nullptr;
    ctx->eval_cap = 0;
    RET(dev_alloc(&ctx->partials, n * PART_DOUBLES));
    RET(dev_alloc(&ctx->scores, n));
    RET(dev_alloc(&ctx->cents, n));
    RET(dev_alloc(&ctx->cand, n * 3));
    ctx->eval_cap = n;
    return SNES_OK;
}

static int ensure_imgs(snes_ctx *ctx, size_t n) {
    if (n <= ctx->img_cap) return SNES_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_imgs);
    cudaFree(ctx->d_imgtm);
    ctx->d_imgtm = nullptr;
    cudaFree(ctx->d_km);
    cudaFree(ctx->best);
    cudaFree(ctx->self_scores);
    cudaFree(ctx->self_xyb);
    cudaFree(ctx->self_partials);
    ctx->self_xyb = nullptr;
    ctx->self_partials = nullptr;
    ctx->d_imgs = nullptr;
    ctx->d_km = nullptr;
    ctx->best = nullptr;
    ctx->self_scores = nullptr;
    ctx->img_cap = 0;
    ctx->cached.clear();
    RET(dev_alloc(&ctx->d_imgs, n));
    RET(dev_alloc(&ctx->d_imgtm, n));
    RET(dev_alloc(&ctx->d_km, n));
    RET(dev_alloc(&ctx->best, n));
    RET(dev_alloc(&ctx->self_scores, n));
    RET(dev_alloc(&ctx->self_xyb, 2 * n * EVAL_XYB_FLOATS));  // [0, n): own palette_map; [n, 2n): prepared base assignment
    RET(dev_alloc(&ctx->self_partials, n * PART_DOUBLES));
    ctx->img_cap = n;
    return SNES_OK;
}

// Validate a batch (same context, same config) and make ctx->d_imgs describe it.
static int bind_images(snes_ctx *ctx, snes_image *const *images, int nimg) {
    if (!ctx || !images || nimg < 1) return fail(SNES_E_INVALID, "batch: no images");
    for (int j = 0; j < nimg; j++) {
        if (!images[j] || images[j]->ctx != ctx) return fail(SNES_E_INVALID, "batch: image belongs to another context");
        if (memcmp(&images[j]->cfg, &images[0]->cfg, sizeof(snes_config)) != 0)
            return fail(SNES_E_INVALID, "batch: all images of a batch must share one config");
    }
    RET(set_device(ctx));
    RET(ensure_imgs(ctx, (size_t)nimg));
    if ((int)ctx->cached.size() == nimg && memcmp(ctx->cached.data(), images, sizeof(snes_image *) * nimg) == 0) return SNES_OK;
    std::vector<ImgDev> h(nimg);
    for (int j = 0; j < nimg; j++) h[j] = images[j]->dev;
    std::vector<ImgTm> htm(nimg);
    for (int j = 0; j < nimg; j++) htm[j] = images[j]->tm;
    CK(cudaMemcpyAsync(ctx->d_imgs, h.data(), sizeof(ImgDev) * nimg, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_imgtm, htm.data(), sizeof(ImgTm) * nimg, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->cached.assign(images, images + nimg);
    return SNES_OK;
}


// ------------------------------------------------------------------------------------------------
// the evaluation pipeline
// ------------------------------------------------------------------------------------------------
struct EvalPlan {
    int nimg = 0;
    int ncand = 1;                    // evaluations per image
    int ovr = -1;                     // palette slot replaced by the candidate colour, -1: none
    const int *d_slots = nullptr;     // device [ncand]: candidate k of every image replaces entry d_slots[k] instead of `ovr` (ovr >= 0 still
                                      // says that the evaluations replace an entry at all)
    const uint8_t *d_cand = nullptr;  // device, required when ovr >= 0: [nimg][cand_stride][3], evaluation (j, k) reads candidate cand_lo + k
    int cand_stride = 0, cand_lo = 0; // candidates per image in d_cand (0: ncand) and the first one this plan evaluates
    bool self = false;                // operate on the images' own palette_map instead of scratch maps
    bool do_assign = false;           // optimize()
    bool do_score = false;            // error()
    uint8_t *d_maps_out = nullptr;    // optional [E][NPIX] device: keep every palette_map
    double *d_scores = nullptr;       // [E] device output of do_score
    const TileMove *d_moves = nullptr;  // [E] device: tile-reassignment candidates (ovr < 0); null = none
    Best *d_best = nullptr;           // with do_score: also the strict-< first minimum of every image's evaluations, as
    int best_idx_base = 0;            // (error, best_idx_base + k), pooled and reduced in one launch (k_pool_argmin)
    bool no_pool = false;             // leave the partial sums unpooled (the caller's finishing kernel pools them)
    bool self_fresh = false;          // every image's palette_map is optimize() of its current state: its coarse pyramid is the
                                      // prepared base assignment's, so error() of the images needs no pyramid of its own
    bool with_self_error = false;     // also error() of every image's own state -> ctx->self_scores and the image's cached
                                      // error, scored inside the first chunk's launch (k_score_v3 only; lib.rs:199, 294)
};

static int launch_scorer(snes_ctx *ctx, const FusedArgs &fa, int ec, const FusedArgs *extra = nullptr, int extra_evals = 0) {
    cudaStream_t st = ctx->stream;
    if (ctx->fused == 3) {
        V3Args va;
        va.f = fa;
        va.nevals = ec;
        va.f2 = extra ? *extra : fa;
        va.nevals2 = extra ? extra_evals : 0;
        va.pair_xb = ctx->pair_xb;
        va.imgtm = ctx->d_imgtm;
        RET(tm_make_evals(&va.tm, fa.from_image ? nullptr : fa.maps, fa.xyb_rm, ec));
        if (extra) RET(tm_make_evals(&va.tm2, extra->from_image ? nullptr : extra->maps, extra->xyb_rm, extra_evals));
        else va.tm2 = va.tm;
        va.counter = ctx->v3_counter;
        va.hscratch = ctx->v3_scratch;
        const int ne = va.nevals + va.nevals2;
        const int items = ((va.pair_xb ? 1 : 3) + (V3_PARTS - 1) * 3) * ne;   // work items: see k_score_v3
        const int grid = items < ctx->nsm * V3_CTAS_PER_SM ? items : ctx->nsm * V3_CTAS_PER_SM;
        CK(cudaMemsetAsync(ctx->v3_counter, 0, 2 * sizeof(int), st));
        if (va.pair_xb) {
            // Scale 0 of channels X and B: one edge-only item per evaluation, in a kernel of its own (score_v3.cuh).  The two
            // kernels are independent and run side by side: k_score_pair (few, long items) is launched first, k_score_v3 (its
            // queue ends in many small items) on the side stream fills the SMs as the pair CTAs retire, so only one tail is
            // exposed -- and for one picture, whose items do not fill the GPU, both grids are resident at once.
            V3Args vp = va;
            vp.counter = ctx->v3_counter + 1;
            vp.hscratch = ctx->v3_scratch + (size_t)ctx->nsm * V3_CTAS_PER_SM * V3_HSCRATCH_FLOATS;
            const int gp = ne < ctx->nsm * V3_PAIR_CTAS_PER_SM ? ne : ctx->nsm * V3_PAIR_CTAS_PER_SM;
            // (profiling: one event pair on the launching stream around fork .. join -- the two kernels overlap, so their own
            // brackets would count the overlap twice)
            prof_begin(ctx, "k_score_pair+k_score_v3");
            CK(cudaEventRecord(ctx->ev_fork, st));
            CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
            k_score_pair<<<gp, V3_THREADS, sizeof(V3PairSmem), st>>>(vp);
            CK(cudaGetLastError());
            k_score_v3<<<grid, V3_THREADS, sizeof(V3Smem), ctx->side>>>(va);
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->ev_join, ctx->side));
            CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
            prof_end(ctx);
            ctx->launches += 2;
        } else {
            LAUNCH(ctx, "k_score_v3", k_score_v3<<<grid, V3_THREADS, sizeof(V3Smem), st>>>(va));
        }
    } else
        LAUNCH(ctx, "k_score_v2", k_score_v2<<<dim3(3, ec), V2_THREADS, sizeof(V2Smem), st>>>(fa));
    return SNES_OK;
}

// Requires bind_images() first.
static int run_plan(snes_ctx *ctx, const snes_config &cfg, const EvalPlan &pl) {
    const int E = pl.nimg * pl.ncand, S = cfg.subpalette_size, CS = cfg.subpalette_count * cfg.subpalette_size;
    cudaStream_t st = ctx->stream;
    RET(ensure_evals(ctx, (size_t)E));
    const int chunk = E < ctx->chunk ? E : ctx->chunk;
    if (pl.do_score || (pl.do_assign && !pl.self && !pl.d_maps_out)) RET(ensure_chunk(ctx, (size_t)chunk));
    const float4 *labtab = cfg.perceptual_palettes ? ctx->labtab : nullptr;

    LAUNCH(ctx, "k_tables", k_tables<<<pl.nimg + (pl.ovr >= 0 ? (E + 255) / 256 : 0), 256, 0, st>>>(ctx->d_imgs, pl.nimg, CS, pl.d_cand,
                                                                          pl.ovr >= 0 ? E : 0, pl.ncand, pl.cand_stride ? pl.cand_stride : pl.ncand,
                                                                          pl.cand_lo, pl.ovr, pl.d_slots, ctx->cents, labtab, ctx->d_fault));

    // error() of the images' own state riding in the candidates' scorer launch: its coarse pyramid and partial sums
    // live in their own buffers; the 3 * nimg extra items join the first chunk
    const bool self_too = pl.with_self_error && ctx->fused == 3 && pl.do_score && !pl.self;
    const bool delta_path = ctx->delta && !cfg.dither && pl.do_assign && pl.do_score && !pl.self && !pl.d_maps_out &&
                            pl.ovr >= 0 && CS <= 255;
    FusedArgs fself;
    const int pyr_gx = pl.nimg < 16 ? 64 : 16;   // few images: one 32x32 region per CTA instead of four in a row
    float *base_xyb = ctx->self_xyb + (size_t)ctx->img_cap * EVAL_XYB_FLOATS;
    const bool share_base = self_too && delta_path && pl.self_fresh;
    if (self_too) {  // the coarse pyramid of the images' own palette_map: input of their error()
        if (!share_base)
            LAUNCH(ctx, "k_pyramid<false>", k_pyramid<false><<<dim3(pyr_gx, pl.nimg), 256, 0, st>>>(ctx->d_imgs, ctx->cents, 1, 0, S, CS, -1, nullptr, 1,
                                                           ctx->self_xyb, 0));
        fself.imgs = ctx->d_imgs;
        fself.cents = ctx->cents;
        fself.ncand = 1;
        fself.e0 = 0;
        fself.S = S;
        fself.CS = CS;
        fself.ovr = -1;
        fself.maps = nullptr;
        fself.from_image = 1;
        fself.gi_fmt = 0;
        fself.xyb_rm = share_base ? base_xyb : ctx->self_xyb;
        fself.partials = ctx->self_partials;
    }

    // no dithering + fused scorer: per-candidate work shrinks to one distance per affected pixel (assign_delta.cuh)
    const bool delta = delta_path;
    if (delta) {
        if (cfg.perceptual_palettes)
            LAUNCH(ctx, "k_assign_prepare<true>", k_assign_prepare<true><<<dim3(64, pl.nimg), 256, 0, st>>>(ctx->d_imgs, S, CS));
        else
            LAUNCH(ctx, "k_assign_prepare<false>", k_assign_prepare<false><<<dim3(64, pl.nimg), 256, 0, st>>>(ctx->d_imgs, S, CS));
        // coarse pyramid of the prepared base assignment (current palette, no candidate): k_assign_pyr copies the blocks
        // a candidate leaves unchanged from it
        LAUNCH(ctx, "k_pyramid<false>", k_pyramid<false><<<dim3(pyr_gx, pl.nimg), 256, 0, st>>>(ctx->d_imgs, ctx->cents, 1, 0, S, CS, -1, nullptr, 2,
                                                       base_xyb, 1));
    }

    for (int e0 = 0; e0 < E; e0 += chunk) {
        const int ec = E - e0 < chunk ? E - e0 : chunk;
        uint8_t *maps = pl.d_maps_out ? pl.d_maps_out + (size_t)e0 * NPIX : ctx->maps;
        if (delta) {
            if (cfg.perceptual_palettes)
                LAUNCH(ctx, "k_assign_pyr<1>", k_assign_pyr<1><<<dim3(4, ec), 256, 0, st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, maps, ctx->xyb_rm, ctx->self_xyb + (size_t)ctx->img_cap * EVAL_XYB_FLOATS));
            else
                LAUNCH(ctx, "k_assign_pyr<0>", k_assign_pyr<0><<<dim3(4, ec), 256, 0, st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, maps, ctx->xyb_rm, ctx->self_xyb + (size_t)ctx->img_cap * EVAL_XYB_FLOATS));
            FusedArgs fa;
            fa.imgs = ctx->d_imgs;
            fa.cents = ctx->cents;
            fa.ncand = pl.ncand;
            fa.e0 = e0;
            fa.S = S;
            fa.CS = CS;
            fa.ovr = pl.ovr;
            fa.maps = maps;
            fa.from_image = 0;
            fa.gi_fmt = 1;
            fa.xyb_rm = ctx->xyb_rm;
            fa.partials = ctx->partials;
            RET(launch_scorer(ctx, fa, ec, (self_too && e0 == 0) ? &fself : nullptr, pl.nimg));
            continue;
        }
        // scratch maps feed only the fused scorer: write global entry indices (no tile_palettes / alpha lookups later)
        int gi = (pl.do_score && !pl.self && !pl.d_maps_out && CS <= 255) ? 1 : 0;
        if (pl.d_moves && pl.do_score) {
            // a tile move changes the tile's subpalette for this evaluation only, so the scorer must not look it up in
            // the image: score from a gi-format scratch map; a palette_map-format copy for the caller is a second pass
            gi = 1;
            if (pl.d_maps_out) {
                uint8_t *outm = maps;
                if (cfg.dither && cfg.perceptual_palettes)
                    LAUNCH(ctx, "k_assign_dither<true>", k_assign_dither<true><<<ec, DITHER_THREADS, dither_smem_bytes(CS, true), st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, outm, 0, 0, pl.d_moves));
                else if (cfg.dither)
                    LAUNCH(ctx, "k_assign_dither<false>", k_assign_dither<false><<<ec, DITHER_THREADS, dither_smem_bytes(CS, false), st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, outm, 0, 0, pl.d_moves));
                else if (cfg.perceptual_palettes)
                    LAUNCH(ctx, "k_assign_lab", k_assign_lab<<<dim3(64, ec), 256, 0, st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, outm, 0, 0, pl.d_moves));
                else
                    LAUNCH(ctx, "k_assign_rgb", k_assign_rgb<<<dim3(64, ec), 256, 0, st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, outm, 0, 0, pl.d_moves));
                maps = ctx->maps;
            }
        }
        if (pl.do_assign) {
            if (cfg.dither && cfg.perceptual_palettes) {
                LAUNCH(ctx, "k_assign_dither<true>", k_assign_dither<true><<<ec, DITHER_THREADS, dither_smem_bytes(CS, true), st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, maps, pl.self, gi, pl.d_moves));
            } else if (cfg.dither) {
                LAUNCH(ctx, "k_assign_dither<false>", k_assign_dither<false><<<ec, DITHER_THREADS, dither_smem_bytes(CS, false), st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, maps, pl.self, gi, pl.d_moves));
            } else if (cfg.perceptual_palettes) {
                LAUNCH(ctx, "k_assign_lab", k_assign_lab<<<dim3(64, ec), 256, 0, st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, maps, pl.self, gi, pl.d_moves));
            } else {
                LAUNCH(ctx, "k_assign_rgb", k_assign_rgb<<<dim3(64, ec), 256, 0, st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, maps, pl.self, gi, pl.d_moves));
            }
        }
        if (!pl.do_score) continue;
        if (gi)
            LAUNCH(ctx, "k_assign_pyr<2>", k_assign_pyr<2><<<dim3(4, ec), 256, 0, st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, maps, ctx->xyb_rm, nullptr));
        else
            LAUNCH(ctx, "k_pyramid<false>", k_pyramid<false><<<dim3(16, ec), 256, 0, st>>>(ctx->d_imgs, ctx->cents, pl.ncand, e0, S, CS, pl.ovr, maps, pl.self,
                                                   ctx->xyb_rm, gi));
        FusedArgs fa;
        fa.imgs = ctx->d_imgs;
        fa.cents = ctx->cents;
        fa.ncand = pl.ncand;
        fa.e0 = e0;
        fa.S = S;
        fa.CS = CS;
        fa.ovr = pl.ovr;
        fa.maps = maps;
        fa.from_image = pl.self;
        fa.gi_fmt = gi;
        fa.xyb_rm = ctx->xyb_rm;
        fa.partials = ctx->partials;
        RET(launch_scorer(ctx, fa, ec, (self_too && e0 == 0) ? &fself : nullptr, pl.nimg));
    }
    // what the last evaluation left in the scratch maps: snes_batch_apply_best_dev / batch_step adopt the accepted candidate's
    // palette_map from there instead of running optimize() again
    ctx->maps_live = pl.do_assign && pl.do_score && !pl.self && !pl.d_maps_out && !pl.d_moves && pl.ovr >= 0 && !pl.d_slots && CS <= 255 &&
                     E <= chunk && pl.cand_lo == 0 && (pl.cand_stride == 0 || pl.cand_stride == pl.ncand);
    ctx->maps_nimg = pl.nimg;
    ctx->maps_ncand = pl.ncand;
    ctx->maps_slot = pl.ovr;
    ctx->maps_cand = pl.d_cand;
    ctx->maps_images = ctx->cached;   // the images bind_images() bound for this plan
    if (pl.no_pool) return SNES_OK;
    if (pl.do_score && pl.d_best && !pl.self) {
        LAUNCH(ctx, "k_pool_argmin", k_pool_argmin<<<pl.nimg, 1024, 0, st>>>(ctx->d_imgs, self_too ? ctx->self_partials : nullptr, ctx->self_scores, ctx->partials,
                                                                  pl.ncand, pl.best_idx_base, pl.d_scores, pl.d_best));
        return SNES_OK;
    }
    if (self_too) {
        LAUNCH(ctx, "k_pool_fused", k_pool_fused<<<(pl.nimg + 3) / 4, 128, 0, st>>>(ctx->self_partials, pl.nimg, ctx->self_scores));
        LAUNCH(ctx, "k_store_cur_err", k_store_cur_err<<<(pl.nimg + 127) / 128, 128, 0, st>>>(ctx->d_imgs, pl.nimg, ctx->self_scores));
    }
    if (pl.do_score) LAUNCH(ctx, "k_pool_fused", k_pool_fused<<<(E + 3) / 4, 128, 0, st>>>(ctx->partials, E, pl.d_scores));
    return SNES_OK;
}

// every image's palette_map is optimize() of its current palette and tile assignment
static bool all_fresh(snes_image *const *images, int nimg) {
    for (int j = 0; j < nimg; j++)
        if (!images[j]->map_fresh) return false;
    return true;
}

// optimize() of every image in the batch (lib.rs:425-501)
static int batch_optimize(snes_ctx *ctx, snes_image *const *images, int nimg) {
    RET(bind_images(ctx, images, nimg));
    EvalPlan pl;
    pl.nimg = nimg;
    pl.self = true;
    pl.do_assign = true;
    RET(run_plan(ctx, images[0]->cfg, pl));
    for (int j = 0; j < nimg; j++) images[j]->map_fresh = true;
    return SNES_OK;
}

// error() of every image in the batch (lib.rs:503-548) -> cur_err of each image (and ctx->self_scores)
static int batch_error(snes_ctx *ctx, snes_image *const *images, int nimg) {
    RET(bind_images(ctx, images, nimg));
    EvalPlan pl;
    pl.nimg = nimg;
    pl.self = true;
    pl.do_score = true;
    pl.d_scores = ctx->self_scores;
    RET(run_plan(ctx, images[0]->cfg, pl));
    LAUNCH(ctx, "k_store_cur_err", k_store_cur_err<<<(nimg + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_imgs, nimg, ctx->self_scores));
    return SNES_OK;
}

// ------------------------------------------------------------------------------------------------
// OptimizedImage
// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int snes_image_new(snes_ctx *ctx, const uint8_t *rgba, int width, int height, const snes_config *cfg, snes_image **out) {
    if (!ctx || !rgba || !cfg || !out) return fail(SNES_E_INVALID, "snes_image_new: NULL argument");
    *out = nullptr;
    // lib.rs:838-840 lets any image with one side of 256 through, after which its fixed 32x32 tile
    // table (lib.rs:58, 565) is wrong; only 256x256 is accepted here.
    if (width != SNES_WIDTH || height != SNES_HEIGHT) return fail(SNES_E_INVALID, "Image must be 256x256");
    const int C = cfg->subpalette_count, S = cfg->subpalette_size;
    if (C < 1 || S < 1 || C > 255 || S > 255 || C * S > MAX_ENTRIES)
        return fail(SNES_E_INVALID, "snes_image_new: need 1 <= subpalette_count*subpalette_size <= 256");
    RET(set_device(ctx));
    snes_image *im = new snes_image();
    im->ctx = ctx;
    im->cfg = *cfg;
    im->cfg.dither = cfg->dither ? 1 : 0;
    im->cfg.perceptual_palettes = cfg->perceptual_palettes ? 1 : 0;
    im->cfg.nes = cfg->nes ? 1 : 0;
    im->cfg.reserved = 0;
    im->alpha.resize(NPIX);
    for (int i = 0; i < NPIX; i++) im->alpha[i] = rgba[4 * i + 3];

    const size_t plane = align_up(sizeof(float) * EVAL_XYB_FLOATS);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += align_up(bytes);
        return o;
    };
    const size_t o_rgba = take(NPIX * 4), o_tp = take(NTILES), o_pal = take(MAX_ENTRIES * 3), o_map = take(NPIX);
    const size_t o_rm = take(plane), o_cm = take(plane), o_mu = take(plane), o_s11 = take(plane), o_ms = take(2 * plane);
    const size_t o_bf = take(sizeof(float2) * NPIX);
    const size_t o_tab = take(sizeof(PalTables)), o_err = take(sizeof(double));
    const size_t o_lab = take(im->cfg.perceptual_palettes ? sizeof(float4) * NPIX : 0);
    const size_t o_alpha = take(NPIX);
    const size_t o_bgi = take(NPIX), o_xi = take(NPIX), o_xk = take(sizeof(int2) * NPIX);
    cudaError_t e = cudaMalloc(&im->slab, off);
    if (e != cudaSuccess) {
        delete im;
        return fail(SNES_E_NOMEM, std::string("snes_image_new: cudaMalloc: ") + cudaGetErrorString(e));
    }
    char *b = (char *)im->slab;
    im->dev.rgba = (const uchar4 *)(b + o_rgba);
    im->dev.tile_pal = (uint8_t *)(b + o_tp);
    im->dev.palette = (uint8_t *)(b + o_pal);
    im->dev.map = (uint8_t *)(b + o_map);
    im->dev.xyb_rm = (const float *)(b + o_rm);
    im->dev.xyb_cm = (const float *)(b + o_cm);
    im->dev.mu1 = (float *)(b + o_mu);
    im->dev.s11 = (float *)(b + o_s11);
    im->dev.ms11 = (float2 *)(b + o_ms);
    im->dev.bfxb = (float2 *)(b + o_bf);
    im->dev.tables = (PalTables *)(b + o_tab);
    im->dev.cur_err = (double *)(b + o_err);
    im->dev.lab = im->cfg.perceptual_palettes ? (const float *)(b + o_lab) : nullptr;
    im->dev.alpha = (const uint8_t *)(b + o_alpha);
    im->dev.base_gi = (uint8_t *)(b + o_bgi);
    im->dev.sec_idx = (uint8_t *)(b + o_xi);
    im->dev.keys = (int2 *)(b + o_xk);

    cudaStream_t st = ctx->stream;
    int rc = SNES_OK;
    auto body = [&]() -> int {
        RET(tm_make_image(&im->tm, im->dev));
        CK(cudaMemsetAsync(im->slab, 0, o_rm, st));  // tile_palettes, palette (Palette::new: all black), palette_map = 0
        CK(cudaMemsetAsync(b + o_tab, 0, sizeof(PalTables), st));
        CK(cudaMemsetAsync(b + o_err, 0, sizeof(double), st));
        CK(cudaMemcpyAsync(b + o_rgba, rgba, NPIX * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b + o_alpha, im->alpha.data(), NPIX, cudaMemcpyHostToDevice, st));
        if (im->cfg.perceptual_palettes) {
            LAUNCH(ctx, "k_image_lab", k_image_lab<<<256, 256, 0, st>>>(im->dev.rgba, (float4 *)(b + o_lab)));
        }
        // source side of SSIMULACRA2, once per image: XYB pyramid, mu1 = blur(i1), s11 = blur(i1*i1)
        snes_image *one[1] = {im};
        RET(bind_images(ctx, one, 1));
        LAUNCH(ctx, "k_pyramid<true>", k_pyramid<true><<<dim3(16, 1), 256, 0, st>>>(ctx->d_imgs, nullptr, 1, 0, 0, 0, -1, nullptr, 0, nullptr, 0));
        for (int s = 0; s < NSCALES; s++) {
            const int d = W >> s, lines = 3 * d;
            LAUNCH(ctx, "k_blur_h", k_blur_h<<<(lines + 127) / 128, 128, kBlurHSmem, st>>>(s, lines, ctx->d_imgs, ctx->hbuf));
            LAUNCH(ctx, "k_blur_v", k_blur_v<<<(lines + 127) / 128, 128, 0, st>>>(s, lines, ctx->d_imgs, ctx->hbuf));
        }
        LAUNCH(ctx, "k_interleave_ms", k_interleave_ms<<<(EVAL_XYB_FLOATS + 255) / 256, 256, 0, st>>>(im->dev.mu1, im->dev.s11, im->dev.ms11, EVAL_XYB_FLOATS));
        LAUNCH(ctx, "k_make_bfxb", k_make_bfxb<<<256, 256, 0, st>>>(im->dev));
        CK(cudaStreamSynchronize(st));
        return SNES_OK;
    };
    rc = body();
    if (rc != SNES_OK) {
        ctx->cached.clear();
        cudaFree(im->slab);
        delete im;
        return rc;
    }
    *out = im;
    return SNES_OK;
}

extern "C" void snes_image_free(snes_image *im) {
    if (!im) return;
    snes_ctx *ctx = im->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->cached.clear();
    ctx->maps_live = false;   // an image created later may reuse this address
    ctx->maps_images.clear();
    cudaFree(im->slab);
    cudaFree(im->km_slab);
    delete im;
}

static int ensure_km(snes_image *im) {
    if (im->km_slab) return SNES_OK;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += align_up(bytes);
        return o;
    };
    const size_t o_pts = take(sizeof(float) * 3 * NPIX), o_asg = take(sizeof(int) * NPIX), o_off = take(sizeof(int) * 260);
    const size_t o_means = take(sizeof(double) * 3 * NTILES), o_tm = take(sizeof(int) * NTILES), o_n = take(sizeof(int));
    const size_t o_cent = take(sizeof(double) * 3 * KM_MAXK), o_st = take(sizeof(int) * 2 * 256);
    CK(cudaMalloc(&im->km_slab, off));
    char *b = (char *)im->km_slab;
    im->km.pts = (float *)(b + o_pts);
    im->km.assign = (int *)(b + o_asg);
    im->km.sub_off = (int *)(b + o_off);
    im->km.means = (double *)(b + o_means);
    im->km.tile_map = (int *)(b + o_tm);
    im->km.nmeans = (int *)(b + o_n);
    im->km.centres = (double *)(b + o_cent);
    im->km.status = (int *)(b + o_st);
    return SNES_OK;
}

static int bind_km(snes_ctx *ctx, snes_image *const *images, int nimg) {
    std::vector<KmScratch> h(nimg);
    for (int j = 0; j < nimg; j++) {
        RET(ensure_km(images[j]));
        h[j] = images[j]->km;
    }
    CK(cudaMemcpyAsync(ctx->d_km, h.data(), sizeof(KmScratch) * nimg, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNES_OK;
}

// recalculate_palette(p) for every subpalette of every image (lib.rs:330-405); first < 0 status wins.
static int batch_recalc(snes_ctx *ctx, snes_image *const *images, int nimg, int only_sub0) {
    const snes_config cfg = images[0]->cfg;
    const int C = only_sub0 ? 1 : cfg.subpalette_count, S = cfg.subpalette_size;
    cudaStream_t st = ctx->stream;
    LAUNCH(ctx, "k_gather_points", k_gather_points<<<nimg, 1024, 0, st>>>(ctx->d_imgs, ctx->d_km, cfg.subpalette_count, cfg.perceptual_palettes));
    for (int j = 0; j < nimg; j++) CK(cudaMemsetAsync(images[j]->km.status, 0xff, sizeof(int) * 2 * 256, st));
    // grid is (image, subpalette) with C as the stride; with only_sub0 the grid covers subpalette 0 only
    // few integer problems (one picture): a thread-block cluster of KM_CLUSTER CTAs per problem (kmeans.cuh); otherwise one CTA each
    if (!cfg.perceptual_palettes && nimg * C * KM_CLUSTER <= 2 * ctx->nsm && !ctx->no_cluster_kmeans)
        LAUNCH(ctx, "k_kmeans_cluster", k_kmeans_cluster<<<nimg * C * KM_CLUSTER, 1024, 0, st>>>(ctx->d_imgs, ctx->d_km, C, S));
    else
        LAUNCH(ctx, "k_kmeans<false>", k_kmeans<false><<<nimg * C, 1024, 0, st>>>(ctx->d_imgs, ctx->d_km, C, S, cfg.perceptual_palettes ? 0 : 1));
    LAUNCH(ctx, "k_centres_to_palette", k_centres_to_palette<<<nimg, 256, 0, st>>>(ctx->d_imgs, ctx->d_km, C, S, cfg.perceptual_palettes, cfg.nes, ctx->labtab, 0));
    std::vector<int> status((size_t)nimg * C);
    for (int j = 0; j < nimg; j++)
        CK(cudaMemcpyAsync(status.data() + (size_t)j * C, images[j]->km.status, sizeof(int) * C, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int v : status)
        if (v != 0) return fail(SNES_E_KMEANS, "cogset Kmeans::new: assertion failed: 2 <= k && k < data.len()");
    return SNES_OK;
}

extern "C" int snes_batch_recalculate_palettes(snes_ctx *ctx, snes_image *const *images, int nimg) {
    RET(bind_images(ctx, images, nimg));
    RET(bind_km(ctx, images, nimg));
    RET(batch_recalc(ctx, images, nimg, 0));
    RET(batch_optimize(ctx, images, nimg));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNES_OK;
}

extern "C" int snes_batch_initialize_tiles(snes_ctx *ctx, snes_image *const *images, int nimg) {
    RET(bind_images(ctx, images, nimg));
    RET(bind_km(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    cudaStream_t st = ctx->stream;
    if (cfg.subpalette_count == 1) {  // lib.rs:80-84
        RET(batch_recalc(ctx, images, nimg, 1));
    } else {
        LAUNCH(ctx, "k_tile_means", k_tile_means<<<nimg, 1024, 0, st>>>(ctx->d_imgs, ctx->d_km, cfg.perceptual_palettes));
        for (int j = 0; j < nimg; j++) CK(cudaMemsetAsync(images[j]->km.status, 0xff, sizeof(int) * 2 * 256, st));
        LAUNCH(ctx, "k_kmeans<true>", k_kmeans<true><<<nimg, 1024, 0, st>>>(ctx->d_imgs, ctx->d_km, cfg.subpalette_count, cfg.subpalette_count, 0));
        LAUNCH(ctx, "k_centres_to_palette", k_centres_to_palette<<<nimg, 256, 0, st>>>(ctx->d_imgs, ctx->d_km, cfg.subpalette_count, cfg.subpalette_size,
                                                  cfg.perceptual_palettes, cfg.nes, ctx->labtab, 1));
        std::vector<int> status(nimg, -1);
        for (int j = 0; j < nimg; j++) CK(cudaMemcpyAsync(&status[j], images[j]->km.status, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int v : status)
            if (v != 0) return fail(SNES_E_KMEANS, "cogset Kmeans::new: assertion failed: 2 <= k && k < data.len()");
    }
    RET(batch_optimize(ctx, images, nimg));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNES_OK;
}

extern "C" int snes_image_initialize_tiles(snes_image *im) {
    if (!im) return fail(SNES_E_INVALID, "image is NULL");
    snes_image *one[1] = {im};
    return snes_batch_initialize_tiles(im->ctx, one, 1);
}

extern "C" int snes_image_recalculate_palettes(snes_image *im) {
    if (!im) return fail(SNES_E_INVALID, "image is NULL");
    snes_image *one[1] = {im};
    return snes_batch_recalculate_palettes(im->ctx, one, 1);
}

extern "C" int snes_batch_optimize(snes_ctx *ctx, snes_image *const *images, int nimg) {
    RET(batch_optimize(ctx, images, nimg));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNES_OK;
}

extern "C" int snes_image_optimize(snes_image *im) {
    if (!im) return fail(SNES_E_INVALID, "image is NULL");
    snes_image *one[1] = {im};
    return snes_batch_optimize(im->ctx, one, 1);
}

extern "C" int snes_batch_error(snes_ctx *ctx, snes_image *const *images, int nimg, double *errors) {
    RET(batch_error(ctx, images, nimg));
    if (errors) CK(cudaMemcpyAsync(errors, ctx->self_scores, sizeof(double) * nimg, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNES_OK;
}

extern "C" int snes_batch_error_dev(snes_ctx *ctx, snes_image *const *images, int nimg, double *d_errors) {
    RET(batch_error(ctx, images, nimg));
    if (d_errors) CK(cudaMemcpyAsync(d_errors, ctx->self_scores, sizeof(double) * nimg, cudaMemcpyDeviceToDevice, ctx->stream));
    return SNES_OK;
}

extern "C" int snes_image_error(snes_image *im, double *err) {
    if (!im || !err) return fail(SNES_E_INVALID, "snes_image_error: NULL argument");
    snes_image *one[1] = {im};
    return snes_batch_error(im->ctx, one, 1, err);
}

extern "C" int snes_image_as_rgba(snes_image *im, uint8_t *out_rgba) {
    if (!im || !out_rgba) return fail(SNES_E_INVALID, "snes_image_as_rgba: NULL argument");
    snes_ctx *ctx = im->ctx;
    RET(set_device(ctx));
    uchar4 *tmp = reinterpret_cast<uchar4 *>(ctx->hbuf);
    LAUNCH(ctx, "k_as_rgba", k_as_rgba<<<256, 256, 0, ctx->stream>>>(im->dev, im->cfg.subpalette_size, tmp));
    CK(cudaMemcpyAsync(out_rgba, tmp, NPIX * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNES_OK;
}

// ---- state accessors ---------------------------------------------------------------------------
static int d2h(snes_image *im, void *dst, const void *src, size_t n) {
    if (!im || !dst) return fail(SNES_E_INVALID, "accessor: NULL argument");
    RET(set_device(im->ctx));
    CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, im->ctx->stream));
    CK(cudaStreamSynchronize(im->ctx->stream));
    return SNES_OK;
}
static int h2d(snes_image *im, void *dst, const void *src, size_t n) {
    if (!im || !src) return fail(SNES_E_INVALID, "accessor: NULL argument");
    RET(set_device(im->ctx));
    CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, im->ctx->stream));
    CK(cudaStreamSynchronize(im->ctx->stream));
    return SNES_OK;
}

extern "C" int snes_image_get_palette(snes_image *im, uint8_t *out) {
    return d2h(im, out, im ? im->dev.palette : nullptr, im ? (size_t)im->cfg.subpalette_count * im->cfg.subpalette_size * 3 : 0);
}
extern "C" int snes_image_set_palette(snes_image *im, const uint8_t *in) {
    if (!im || !in) return fail(SNES_E_INVALID, "accessor: NULL argument");
    const size_t n = (size_t)im->cfg.subpalette_count * im->cfg.subpalette_size * 3;
    // SnesColor values are 5-bit; 32 can arise from round(v/8) (lib.rs:396-400) and is kept
    for (size_t i = 0; i < n; i++)
        if (in[i] > 32) return fail(SNES_E_INVALID, "snes_image_set_palette: colour component > 32");
    im->map_fresh = false;
    return h2d(im, im->dev.palette, in, n);
}
extern "C" int snes_image_get_tile_palettes(snes_image *im, uint8_t *out) { return d2h(im, out, im ? im->dev.tile_pal : nullptr, NTILES); }
extern "C" int snes_image_set_tile_palettes(snes_image *im, const uint8_t *in) {
    if (!im || !in) return fail(SNES_E_INVALID, "accessor: NULL argument");
    for (int i = 0; i < NTILES; i++)
        if (in[i] >= im->cfg.subpalette_count) return fail(SNES_E_INVALID, "snes_image_set_tile_palettes: index >= subpalette_count");
    im->map_fresh = false;
    return h2d(im, im->dev.tile_pal, in, NTILES);
}
extern "C" int snes_image_get_palette_map(snes_image *im, uint8_t *out) { return d2h(im, out, im ? im->dev.map : nullptr, NPIX); }
extern "C" int snes_image_set_palette_map(snes_image *im, const uint8_t *in) {
    if (!im || !in) return fail(SNES_E_INVALID, "accessor: NULL argument");
    for (int i = 0; i < NPIX; i++)
        if (in[i] >= im->cfg.subpalette_size) return fail(SNES_E_INVALID, "snes_image_set_palette_map: index >= subpalette_size");
    im->map_fresh = false;
    return h2d(im, im->dev.map, in, NPIX);
}

// FNV-1a (64-bit) over palette, tile_palettes and palette_map: what replicas of an image on different ranks must agree on.
extern "C" int snes_image_state_checksum(snes_image *im, uint64_t *out) {
    if (!im || !out) return fail(SNES_E_INVALID, "snes_image_state_checksum: NULL argument");
    const size_t np = (size_t)im->cfg.subpalette_count * im->cfg.subpalette_size * 3;
    std::vector<uint8_t> buf(np + NTILES + NPIX);
    RET(set_device(im->ctx));
    cudaStream_t st = im->ctx->stream;
    CK(cudaMemcpyAsync(buf.data(), im->dev.palette, np, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(buf.data() + np, im->dev.tile_pal, NTILES, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(buf.data() + np + NTILES, im->dev.map, NPIX, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint8_t b : buf) h = (h ^ b) * 0x100000001b3ull;
    *out = h;
    return SNES_OK;
}

// lib.rs:579-625 + serde_json Value::to_string(): compact, object keys in BTreeMap (sorted) order.
extern "C" int snes_image_as_json(snes_image *im, char *buf, size_t cap, size_t *len) {
    if (!im || !len) return fail(SNES_E_INVALID, "snes_image_as_json: NULL argument");
    const int C = im->cfg.subpalette_count, S = im->cfg.subpalette_size;
    std::vector<uint8_t> pal((size_t)C * S * 3), tp(NTILES), map(NPIX);
    RET(snes_image_get_palette(im, pal.data()));
    RET(snes_image_get_tile_palettes(im, tp.data()));
    RET(snes_image_get_palette_map(im, map.data()));
    std::string s;
    s.reserve(NPIX * 3 + 16384);
    char num[16];
    s += "{\"palette\":[";
    for (int p = 0; p < C; p++)
        for (int i = 0; i < 16; i++) {
            unsigned v = 0;
            if (i != 0 && i <= S) {
                const uint8_t *c = &pal[3 * (size_t)(p * S + i - 1)];
                v = (unsigned)c[0] + ((unsigned)c[1] << 5) + ((unsigned)c[2] << 10);  // as_u16, lib.rs:679-681
                v &= 0xffffu;
            }
            snprintf(num, sizeof num, "%u", v);
            if (p || i) s += ',';
            s += num;
        }
    s += "],\"tile_palettes\":[";
    for (int t = 0; t < NTILES; t++) {
        snprintf(num, sizeof num, "%u", (unsigned)tp[t]);
        if (t) s += ',';
        s += num;
    }
    s += "],\"tiles\":[";
    for (int ty = 0; ty < 32; ty++)
        for (int tx = 0; tx < 32; tx++) {
            if (ty || tx) s += ',';
            s += '[';
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    const int index = (ty * 8 + y) * W + tx * 8 + x;
                    const unsigned v = im->alpha[index] == 0 ? 0u : (unsigned)map[index] + 1u;
                    snprintf(num, sizeof num, "%u", v);
                    if (y || x) s += ',';
                    s += num;
                }
            s += ']';
        }
    s += "]}";
    *len = s.size();
    if (buf && cap > 0) {
        const size_t n = s.size() < cap ? s.size() : cap;
        memcpy(buf, s.data(), n);
        if (s.size() < cap) buf[s.size()] = '\0';
    }
    return SNES_OK;
}

// ------------------------------------------------------------------------------------------------
// batched candidate evaluation
// ------------------------------------------------------------------------------------------------
static int check_slot(const snes_config &cfg, int palette, int index) {
    if (palette < 0 || palette >= cfg.subpalette_count || index < 0 || index >= cfg.subpalette_size)
        return fail(SNES_E_INVALID, "palette/index out of range");
    return SNES_OK;
}

static int eval_candidates_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index, const uint8_t *d_cand,
                               int ncand, int cand_idx_base, double *d_scores, snes_best *d_best, bool with_error,
                               int cand_stride = 0, int cand_lo = 0) {
    RET(bind_images(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    RET(check_slot(cfg, palette, index));
    if (!d_cand || ncand < 0) return fail(SNES_E_INVALID, "snes_batch_eval_candidates_dev: no candidates");
    if (cand_stride && (cand_lo < 0 || cand_lo + ncand > cand_stride)) return fail(SNES_E_INVALID, "candidate slice outside the list");
    if (ncand == 0) {
        // an empty slice (more ranks than candidates): nothing to evaluate; the records say so (idx -1 never wins the merge)
        if (with_error) RET(batch_error(ctx, images, nimg));
        if (d_best) LAUNCH(ctx, "k_no_best", k_no_best<<<(nimg + 127) / 128, 128, 0, ctx->stream>>>(reinterpret_cast<Best *>(d_best), nimg));
        return SNES_OK;
    }
    RET(ensure_evals(ctx, (size_t)nimg * ncand));
    if (with_error && ctx->fused != 3) RET(batch_error(ctx, images, nimg));
    EvalPlan pl;
    pl.nimg = nimg;
    pl.ncand = ncand;
    pl.ovr = palette * cfg.subpalette_size + index;
    pl.d_cand = d_cand;
    pl.cand_stride = cand_stride;
    pl.cand_lo = cand_lo;
    pl.do_assign = pl.do_score = true;
    pl.d_scores = d_scores ? d_scores : ctx->scores;
    pl.with_self_error = with_error && ctx->fused == 3;
    pl.self_fresh = all_fresh(images, nimg);
    pl.d_best = reinterpret_cast<Best *>(d_best);
    pl.best_idx_base = cand_idx_base;
    RET(run_plan(ctx, cfg, pl));
    return SNES_OK;
}

extern "C" int snes_batch_eval_candidates_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                              const uint8_t *d_cand, int ncand, int cand_idx_base, double *d_scores,
                                              snes_best *d_best) {
    return eval_candidates_dev(ctx, images, nimg, palette, index, d_cand, ncand, cand_idx_base, d_scores, d_best, false);
}

extern "C" int snes_batch_error_eval_candidates_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                                    const uint8_t *d_cand, int ncand, int cand_idx_base, double *d_scores,
                                                    snes_best *d_best) {
    return eval_candidates_dev(ctx, images, nimg, palette, index, d_cand, ncand, cand_idx_base, d_scores, d_best, true);
}

extern "C" int snes_batch_error_eval_candidates_slice_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                                          const uint8_t *d_cand_all, int ncand_all, int cand_lo, int ncand,
                                                          double *d_scores, snes_best *d_best) {
    if (ncand_all < 1) return fail(SNES_E_INVALID, "snes_batch_error_eval_candidates_slice_dev: no candidates");
    return eval_candidates_dev(ctx, images, nimg, palette, index, d_cand_all, ncand, cand_lo, d_scores, d_best, true, ncand_all, cand_lo);
}

extern "C" int snes_batch_eval_candidates(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                          const uint8_t *cand, int ncand, double *scores, uint8_t *maps, snes_best *best) {
    RET(bind_images(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    RET(check_slot(cfg, palette, index));
    if (!cand || ncand < 1) return fail(SNES_E_INVALID, "snes_batch_eval_candidates: no candidates");
    const size_t E = (size_t)nimg * ncand;
    for (size_t i = 0; i < E * 3; i++)
        if (cand[i] > 32) return fail(SNES_E_INVALID, "snes_batch_eval_candidates: colour component > 32");
    RET(ensure_evals(ctx, E));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->cand, cand, E * 3, cudaMemcpyHostToDevice, st));
    uint8_t *d_maps = nullptr;
    if (maps) CK(cudaMalloc((void **)&d_maps, E * NPIX));
    EvalPlan pl;
    pl.nimg = nimg;
    pl.ncand = ncand;
    pl.ovr = palette * cfg.subpalette_size + index;
    pl.d_cand = ctx->cand;
    pl.do_assign = pl.do_score = true;
    pl.d_maps_out = d_maps;
    pl.d_scores = ctx->scores;
    int rc = run_plan(ctx, cfg, pl);
    if (rc == SNES_OK && best) {
        rc = [&]() -> int {
            LAUNCH(ctx, "k_argmin", k_argmin<<<nimg, 128, 0, st>>>(ctx->scores, ncand, 0, ctx->best));
            return SNES_OK;
        }();
    }
    auto copy_out = [&]() -> int {
        if (scores) CK(cudaMemcpyAsync(scores, ctx->scores, sizeof(double) * E, cudaMemcpyDeviceToHost, st));
        if (best) CK(cudaMemcpyAsync(best, ctx->best, sizeof(Best) * nimg, cudaMemcpyDeviceToHost, st));
        if (maps) CK(cudaMemcpyAsync(maps, d_maps, E * NPIX, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return SNES_OK;
    };
    if (rc == SNES_OK) rc = copy_out();
    if (d_maps) cudaFree(d_maps);
    return rc;
}

// ---- tile reassignment as evaluated candidates (TODO.md:36-37) ------------------------------
static int upload_moves(snes_ctx *ctx, const snes_config &cfg, const int32_t *moves, size_t n, TileMove **d_out) {
    if (!moves || n < 1) return fail(SNES_E_INVALID, "tile moves: no candidates");
    for (size_t i = 0; i < n; i++)
        if (moves[2 * i] < 0 || moves[2 * i] >= NTILES || moves[2 * i + 1] < 0 || moves[2 * i + 1] >= cfg.subpalette_count)
            return fail(SNES_E_INVALID, "tile moves: tile must be in 0..1023 and subpalette below subpalette_count");
    TileMove *d = nullptr;
    CK(cudaMalloc((void **)&d, sizeof(TileMove) * n));
    cudaError_t e = cudaMemcpyAsync(d, moves, sizeof(TileMove) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(d);
        return fail(SNES_E_CUDA, cudaGetErrorString(e));
    }
    *d_out = d;
    return SNES_OK;
}

static int eval_tile_moves(snes_ctx *ctx, snes_image *const *images, int nimg, const snes_config &cfg, const TileMove *d_moves,
                           int nmoves, uint8_t *d_maps) {
    if (cfg.subpalette_count * cfg.subpalette_size > 255)
        return fail(SNES_E_INVALID, "tile moves need subpalette_count * subpalette_size <= 255");
    EvalPlan pl;
    pl.nimg = nimg;
    pl.ncand = nmoves;
    pl.ovr = -1;
    pl.d_moves = d_moves;
    pl.do_assign = pl.do_score = true;
    pl.d_maps_out = d_maps;
    pl.d_scores = ctx->scores;
    RET(run_plan(ctx, cfg, pl));
    LAUNCH(ctx, "k_argmin", k_argmin<<<nimg, 128, 0, ctx->stream>>>(ctx->scores, nmoves, 0, ctx->best));
    return SNES_OK;
}

extern "C" int snes_batch_eval_tile_moves(snes_ctx *ctx, snes_image *const *images, int nimg, const int32_t *moves, int nmoves,
                                          double *scores, uint8_t *maps, snes_best *best) {
    RET(bind_images(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    const size_t E = (size_t)nimg * (nmoves > 0 ? nmoves : 0);
    TileMove *d_moves = nullptr;
    RET(upload_moves(ctx, cfg, moves, E, &d_moves));
    uint8_t *d_maps = nullptr;
    cudaStream_t st = ctx->stream;
    int rc = [&]() -> int {
        RET(ensure_evals(ctx, E));
        if (maps) CK(cudaMalloc((void **)&d_maps, E * NPIX));
        RET(eval_tile_moves(ctx, images, nimg, cfg, d_moves, nmoves, d_maps));
        if (scores) CK(cudaMemcpyAsync(scores, ctx->scores, sizeof(double) * E, cudaMemcpyDeviceToHost, st));
        if (best) CK(cudaMemcpyAsync(best, ctx->best, sizeof(Best) * nimg, cudaMemcpyDeviceToHost, st));
        if (maps) CK(cudaMemcpyAsync(maps, d_maps, E * NPIX, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return SNES_OK;
    }();
    cudaStreamSynchronize(st);
    cudaFree(d_moves);
    if (d_maps) cudaFree(d_maps);
    return rc;
}

extern "C" int snes_batch_step_tile_moves(snes_ctx *ctx, snes_image *const *images, int nimg, const int32_t *moves, int nmoves,
                                          snes_best *best, uint8_t *applied) {
    RET(bind_images(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    const size_t E = (size_t)nimg * (nmoves > 0 ? nmoves : 0);
    TileMove *d_moves = nullptr;
    RET(upload_moves(ctx, cfg, moves, E, &d_moves));
    uint8_t *d_applied = nullptr;
    cudaStream_t st = ctx->stream;
    int rc = [&]() -> int {
        RET(ensure_evals(ctx, E));
        CK(cudaMalloc((void **)&d_applied, nimg));
        RET(batch_error(ctx, images, nimg));   // the error the best move has to beat
        RET(eval_tile_moves(ctx, images, nimg, cfg, d_moves, nmoves, nullptr));
        LAUNCH(ctx, "k_apply_tile_move", k_apply_tile_move<<<(nimg + 127) / 128, 128, 0, st>>>(ctx->d_imgs, nimg, d_moves, nmoves, ctx->best, d_applied));
        RET(batch_optimize(ctx, images, nimg));
        if (best) CK(cudaMemcpyAsync(best, ctx->best, sizeof(Best) * nimg, cudaMemcpyDeviceToHost, st));
        if (applied) CK(cudaMemcpyAsync(applied, d_applied, nimg, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return SNES_OK;
    }();
    cudaStreamSynchronize(st);
    cudaFree(d_moves);
    if (d_applied) cudaFree(d_applied);
    return rc;
}

// Accept step + the optimize() that follows it (lib.rs:216-219 + 236-237, 280-281, 324-325).  When the candidates were
// just evaluated by this context in full (every candidate of every image, maps still in the scratch buffer) and the images'
// palette_maps were current, optimize() of an image that accepted a candidate is that candidate's map (k_adopt_map) and
// optimize() of one that did not changes nothing; otherwise optimize() runs.
static int ensure_ints(snes_ctx *ctx, size_t n);
static int apply_and_optimize(snes_ctx *ctx, snes_image *const *images, int nimg, int slot, const uint8_t *d_cand_all, int ncand_all,
                              const Best *d_best, int force) {
    cudaStream_t st = ctx->stream;
    const snes_config cfg = images[0]->cfg;
    bool adopt = ctx->maps_live && ctx->maps_nimg == nimg && ctx->maps_ncand == ncand_all && ctx->maps_slot == slot &&
                 ctx->maps_cand == d_cand_all && (int)ctx->maps_images.size() == nimg && all_fresh(images, nimg);
    for (int j = 0; adopt && j < nimg; j++) adopt = ctx->maps_images[j] == images[j];
    ctx->maps_live = false;
    if (adopt) {
        RET(ensure_ints(ctx, (size_t)nimg));
        LAUNCH(ctx, "k_apply_best", k_apply_best<<<(nimg + 127) / 128, 128, 0, st>>>(ctx->d_imgs, nimg, slot, d_cand_all, ncand_all, d_best, force, ctx->d_ints));
        LAUNCH(ctx, "k_adopt_map", k_adopt_map<<<dim3(64, nimg), 256, 0, st>>>(ctx->d_imgs, ctx->d_ints, ctx->maps, ncand_all, cfg.subpalette_size));
        return SNES_OK;
    }
    LAUNCH(ctx, "k_apply_best", k_apply_best<<<(nimg + 127) / 128, 128, 0, st>>>(ctx->d_imgs, nimg, slot, d_cand_all, ncand_all, d_best, force));
    return batch_optimize(ctx, images, nimg);
}

extern "C" int snes_merge_best_dev(snes_ctx *ctx, const snes_best *d_gathered, int nranks, int rank_stride, int nimg, snes_best *d_out) {
    if (!ctx || !d_gathered || !d_out || nranks < 1 || nimg < 1 || rank_stride < nimg) return fail(SNES_E_INVALID, "snes_merge_best_dev: bad argument");
    RET(set_device(ctx));
    LAUNCH(ctx, "k_merge_best", k_merge_best<<<(nimg + 127) / 128, 128, 0, ctx->stream>>>(reinterpret_cast<const Best *>(d_gathered), nranks, rank_stride, nimg,
                                                           reinterpret_cast<Best *>(d_out)));
    return SNES_OK;
}

extern "C" int snes_batch_apply_best_dev(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                         const uint8_t *d_cand_all, int ncand_all, const snes_best *d_best) {
    RET(bind_images(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    RET(check_slot(cfg, palette, index));
    if (!d_cand_all || !d_best || ncand_all < 1) return fail(SNES_E_INVALID, "snes_batch_apply_best_dev: NULL argument");
    return apply_and_optimize(ctx, images, nimg, palette * cfg.subpalette_size + index, d_cand_all, ncand_all, reinterpret_cast<const Best *>(d_best),
                              cfg.nes ? 1 : 0);
}

// One optimize_palette_entry_* call for every image of the batch.
//   mode 0: random (explicit candidates, lib.rs:191-240), 1: NES (lib.rs:242-284), 2: channel (lib.rs:286-328)
static int batch_step(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index, int mode, int channel,
                      const uint8_t *cand, int ncand, snes_best *best, double *errors_after) {
    RET(bind_images(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    RET(check_slot(cfg, palette, index));
    const int slot = palette * cfg.subpalette_size + index;
    if (mode == 1) ncand = NES_COUNT;
    if (mode == 2) ncand = 32;
    if (mode == 2 && (channel < 0 || channel > 2)) return fail(SNES_E_INVALID, "channel out of range");
    if (mode == 0 && (!cand || ncand < 1)) return fail(SNES_E_INVALID, "no candidates");
    const size_t E = (size_t)nimg * ncand;
    RET(ensure_evals(ctx, E));
    cudaStream_t st = ctx->stream;
    if (mode == 0) {
        for (size_t i = 0; i < E * 3; i++)
            if (cand[i] > 32) return fail(SNES_E_INVALID, "colour component > 32");
        CK(cudaMemcpyAsync(ctx->cand, cand, E * 3, cudaMemcpyHostToDevice, st));
    } else {
        LAUNCH(ctx, "k_make_cands", k_make_cands<<<nimg, 64, 0, st>>>(ctx->d_imgs, slot, mode == 1 ? -1 : channel, ctx->cand, ncand));
    }
    // best_error = self.error() (lib.rs:199, 294): inside the candidates' scorer launch with k_score_v3, else on its own
    if (mode != 1 && ctx->fused != 3) RET(batch_error(ctx, images, nimg));
    EvalPlan pl;
    pl.nimg = nimg;
    pl.ncand = ncand;
    pl.ovr = slot;
    pl.d_cand = ctx->cand;
    pl.do_assign = pl.do_score = true;
    pl.d_scores = ctx->scores;
    pl.with_self_error = mode != 1 && ctx->fused == 3;
    pl.self_fresh = all_fresh(images, nimg);
    pl.d_best = ctx->best;
    RET(run_plan(ctx, cfg, pl));
    RET(apply_and_optimize(ctx, images, nimg, slot, ctx->cand, ncand, ctx->best, mode == 1));
    if (best) CK(cudaMemcpyAsync(best, ctx->best, sizeof(Best) * nimg, cudaMemcpyDeviceToHost, st));
    if (errors_after) {
        RET(batch_error(ctx, images, nimg));
        CK(cudaMemcpyAsync(errors_after, ctx->self_scores, sizeof(double) * nimg, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return SNES_OK;
}

// ---- one optimiser step sharded over the GPUs of a box, host buffers in and out -------------------
// The collective in the middle (an all-gather of 16 bytes per image and rank) belongs to the caller's process group;
// the library provides the two halves around it and keeps the candidate list on the device in between.
extern "C" int snes_batch_step_random_shard_begin(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                                  const uint8_t *cand, int ncand_all, int cand_lo, int ncand,
                                                  snes_best *d_best_local) {
    RET(bind_images(ctx, images, nimg));
    if (!cand || ncand_all < 1 || !d_best_local) return fail(SNES_E_INVALID, "snes_batch_step_random_shard_begin: bad argument");
    const size_t E = (size_t)nimg * ncand_all;
    for (size_t i = 0; i < E * 3; i++)
        if (cand[i] > 32) return fail(SNES_E_INVALID, "colour component > 32");
    RET(ensure_evals(ctx, E));
    CK(cudaMemcpyAsync(ctx->cand, cand, E * 3, cudaMemcpyHostToDevice, ctx->stream));
    ctx->shard_ncand_all = ncand_all;
    return eval_candidates_dev(ctx, images, nimg, palette, index, ctx->cand, ncand, cand_lo, nullptr, d_best_local, true, ncand_all, cand_lo);
}

extern "C" int snes_batch_step_random_shard_end(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                                const snes_best *d_gathered, int nranks, int rank_stride, snes_best *best,
                                                double *errors_after) {
    RET(bind_images(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    RET(check_slot(cfg, palette, index));
    if (!d_gathered || nranks < 1 || rank_stride < nimg || ctx->shard_ncand_all < 1)
        return fail(SNES_E_INVALID, "snes_batch_step_random_shard_end: bad argument (or no _begin before it)");
    cudaStream_t st = ctx->stream;
    LAUNCH(ctx, "k_merge_best", k_merge_best<<<(nimg + 127) / 128, 128, 0, st>>>(reinterpret_cast<const Best *>(d_gathered), nranks, rank_stride, nimg, ctx->best));
    RET(apply_and_optimize(ctx, images, nimg, palette * cfg.subpalette_size + index, ctx->cand, ctx->shard_ncand_all, ctx->best, cfg.nes ? 1 : 0));
    if (best) CK(cudaMemcpyAsync(best, ctx->best, sizeof(Best) * nimg, cudaMemcpyDeviceToHost, st));
    if (errors_after) {
        RET(batch_error(ctx, images, nimg));
        CK(cudaMemcpyAsync(errors_after, ctx->self_scores, sizeof(double) * nimg, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return SNES_OK;
}

// ---- the sharded step with the collective inside the library -------------------------------------------------------------
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy a host process already holds -- torch's -- or the system's), so
// the library has no link-time dependency on it and loads on a box without NCCL.  The unique id travels through the ABI: the
// host creates it on one rank and hands it to the others by whatever means it has (a Rust host: a file, a socket, MPI).
struct snes_nccl_id {
    char internal[128];
};
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(snes_nccl_id *) = nullptr;
    int (*CommInitRank)(void **, int, snes_nccl_id, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (int (*)(snes_nccl_id *))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void **, int, snes_nccl_id, int))dlsym(h, "ncclCommInitRank");
            api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
            api.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
            api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy) api.lib = h;
        }
    }
    return api.lib ? &api : nullptr;
}
static int nccl_fail(NcclApi *api, const char *what, int rc) {
    return fail(SNES_E_CUDA, std::string(what) + ": " + (api && api->GetErrorString ? api->GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}

extern "C" int snes_comm_unique_id(uint8_t *out128) {
    if (!out128) return fail(SNES_E_INVALID, "snes_comm_unique_id: NULL argument");
    NcclApi *api = nccl_api();
    if (!api) return fail(SNES_E_CUDA, "libnccl.so.2 could not be loaded");
    snes_nccl_id id;
    const int rc = api->GetUniqueId(&id);
    if (rc != 0) return nccl_fail(api, "ncclGetUniqueId", rc);
    memcpy(out128, id.internal, 128);
    return SNES_OK;
}

extern "C" int snes_ctx_comm_init(snes_ctx *ctx, const uint8_t *id128, int rank, int world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(SNES_E_INVALID, "snes_ctx_comm_init: bad argument");
    NcclApi *api = nccl_api();
    if (!api) return fail(SNES_E_CUDA, "libnccl.so.2 could not be loaded");
    RET(set_device(ctx));
    snes_ctx_comm_destroy(ctx);
    snes_nccl_id id;
    memcpy(id.internal, id128, 128);
    void *comm = nullptr;
    const int rc = api->CommInitRank(&comm, world, id, rank);
    if (rc != 0) return nccl_fail(api, "ncclCommInitRank", rc);
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return SNES_OK;
}

extern "C" int snes_ctx_comm_destroy(snes_ctx *ctx) {
    if (!ctx) return fail(SNES_E_INVALID, "ctx is NULL");
    if (ctx->nccl_comm) {
        NcclApi *api = nccl_api();
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (api) api->CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    return SNES_OK;
}

// driver.plan_shards in C: image groups x candidate slices.  As many groups as divide the world and do not exceed the number of
// images; the ranks of a group slice the candidates.
extern "C" int snes_dist_plan(int nimg_total, int rank, int world, int *img_lo, int *img_hi, int *cand_ranks, int *slice, int *slots) {
    if (nimg_total < 1 || world < 1 || rank < 0 || rank >= world) return fail(SNES_E_INVALID, "snes_dist_plan: bad argument");
    int groups = 1;
    for (int d = 1; d <= world; d++)
        if (world % d == 0 && d <= nimg_total) groups = d;
    const int cr = world / groups, g = rank / cr;
    if (img_lo) *img_lo = (int)(((long long)g * nimg_total) / groups);
    if (img_hi) *img_hi = (int)(((long long)(g + 1) * nimg_total) / groups);
    if (cand_ranks) *cand_ranks = cr;
    if (slice) *slice = rank % cr;
    if (slots) *slots = (nimg_total + groups - 1) / groups;
    return SNES_OK;
}

extern "C" int snes_dist_step_random(snes_ctx *ctx, snes_image *const *images, int nimg_local, int nimg_total, int palette, int index,
                                     const uint8_t *cand, int ncand, snes_best *best, snes_best *all_best) {
    RET(bind_images(ctx, images, nimg_local));
    const snes_config cfg = images[0]->cfg;
    RET(check_slot(cfg, palette, index));
    if (!cand || ncand < 1) return fail(SNES_E_INVALID, "snes_dist_step_random: no candidates");
    const int world = ctx->comm_world, rank = ctx->comm_rank;
    if (world > 1 && !ctx->nccl_comm) return fail(SNES_E_INVALID, "snes_dist_step_random: no communicator (snes_ctx_comm_init)");
    int lo_img, hi_img, cr, slice, slots;
    RET(snes_dist_plan(nimg_total, rank, world, &lo_img, &hi_img, &cr, &slice, &slots));
    if (hi_img - lo_img != nimg_local) return fail(SNES_E_INVALID, "snes_dist_step_random: this rank must hold images [" + std::to_string(lo_img) + ", " +
                                                                       std::to_string(hi_img) + ") of the job (snes_dist_plan)");
    const size_t E = (size_t)nimg_local * ncand;
    for (size_t i = 0; i < E * 3; i++)
        if (cand[i] > 32) return fail(SNES_E_INVALID, "colour component > 32");
    RET(ensure_evals(ctx, E));
    if ((size_t)world * slots > ctx->dist_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_send);
        cudaFree(ctx->d_gather);
        ctx->d_send = ctx->d_gather = nullptr;
        ctx->dist_cap = 0;
        RET(dev_alloc(&ctx->d_send, (size_t)slots));
        RET(dev_alloc(&ctx->d_gather, (size_t)world * slots));
        ctx->dist_cap = (size_t)world * slots;
    }
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->cand, cand, E * 3, cudaMemcpyHostToDevice, st));
    LAUNCH(ctx, "k_no_best", k_no_best<<<(slots + 127) / 128, 128, 0, st>>>(ctx->d_send, slots));   // padding slots of a smaller group
    const int lo = (int)(((long long)slice * ncand) / cr), hi = (int)(((long long)(slice + 1) * ncand) / cr);
    RET(eval_candidates_dev(ctx, images, nimg_local, palette, index, ctx->cand, hi - lo, lo, nullptr, reinterpret_cast<snes_best *>(ctx->d_send), true, ncand, lo));
    const Best *mine = ctx->d_send;
    int nranks = 1, stride = slots;
    if (world > 1) {
        NcclApi *api = nccl_api();
        const int rc = api->AllGather(ctx->d_send, ctx->d_gather, (size_t)slots * sizeof(Best), /* ncclChar */ 0, ctx->nccl_comm, st);
        if (rc != 0) return nccl_fail(api, "ncclAllGather", rc);
        mine = ctx->d_gather + (size_t)(rank - slice) * slots;   // the records of this rank's group: cr consecutive ranks
        nranks = cr;
    }
    LAUNCH(ctx, "k_merge_best", k_merge_best<<<(nimg_local + 127) / 128, 128, 0, st>>>(mine, nranks, stride, nimg_local, ctx->best));
    RET(apply_and_optimize(ctx, images, nimg_local, palette * cfg.subpalette_size + index, ctx->cand, ncand, ctx->best, cfg.nes ? 1 : 0));
    if (best) CK(cudaMemcpyAsync(best, ctx->best, sizeof(Best) * nimg_local, cudaMemcpyDeviceToHost, st));
    if (all_best) CK(cudaMemcpyAsync(all_best, world > 1 ? ctx->d_gather : ctx->d_send, sizeof(Best) * (size_t)world * slots, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SNES_OK;
}

extern "C" int snes_batch_step_random(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index,
                                      const uint8_t *cand, int ncand, snes_best *best, double *errors_after) {
    return batch_step(ctx, images, nimg, palette, index, 0, 0, cand, ncand, best, errors_after);
}
extern "C" int snes_batch_step_nes(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index, snes_best *best,
                                   double *errors_after) {
    return batch_step(ctx, images, nimg, palette, index, 1, 0, nullptr, 0, best, errors_after);
}
extern "C" int snes_batch_step_channel(snes_ctx *ctx, snes_image *const *images, int nimg, int palette, int index, int channel,
                                       snes_best *best, double *errors_after) {
    return batch_step(ctx, images, nimg, palette, index, 2, channel, nullptr, 0, best, errors_after);
}

// ---- several palette entries in one launch sequence (TODO.md:33-35) -----------------------------
// Evaluation (j, s * ncand + k) = image j with entry steps[s] replaced by candidate k of step s.  Every kernel reads the
// replaced entry per evaluation (CandEntry::slot), and k_assign_prepare's records do not depend on it, so the candidates of
// a whole sweep -- or of the next few iterations of one picture -- share one k_tables / prepare / assign / score sequence.
static int ensure_ints(snes_ctx *ctx, size_t n) {
    if (n <= ctx->ints_cap) return SNES_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_ints);
    ctx->d_ints = nullptr;
    ctx->ints_cap = 0;
    RET(dev_alloc(&ctx->d_ints, n));
    ctx->ints_cap = n;
    return SNES_OK;
}
static int ensure_best_m(snes_ctx *ctx, size_t n) {
    if (n <= ctx->best_m_cap) return SNES_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->best_m);
    ctx->best_m = nullptr;
    ctx->best_m_cap = 0;
    RET(dev_alloc(&ctx->best_m, n));
    ctx->best_m_cap = n;
    return SNES_OK;
}

struct MultiLayout {   // offsets into ctx->d_ints
    int *slots, *step_slot, *step_channel, *consumed, *chosen;
};

// mode 0: explicit candidates cand[nimg][nsteps][ncand][3] (host); 1: the 56 NES colours; 2: the 32 values of steps[s].channel.
// Leaves scores in ctx->scores ([nimg][nsteps][ncand]) and the per-step first minima in ctx->best_m ([nimg][nsteps]).
static int eval_multi(snes_ctx *ctx, snes_image *const *images, int nimg, int mode, const snes_step *steps, int nsteps, const uint8_t *cand,
                      int &ncand, bool with_error, MultiLayout &lay, bool finish_later = false) {
    RET(bind_images(ctx, images, nimg));
    const snes_config cfg = images[0]->cfg;
    if (!steps || nsteps < 1 || nsteps > 4096) return fail(SNES_E_INVALID, "multi-entry call: need 1..4096 steps");
    if (mode < 0 || mode > 2) return fail(SNES_E_INVALID, "multi-entry call: mode must be 0 (random), 1 (NES) or 2 (channel)");
    if (mode == 1) ncand = NES_COUNT;
    if (mode == 2) ncand = 32;
    if (mode == 0 && (!cand || ncand < 1)) return fail(SNES_E_INVALID, "no candidates");
    for (int s = 0; s < nsteps; s++) {
        RET(check_slot(cfg, steps[s].palette, steps[s].index));
        if (mode == 2 && (steps[s].channel < 0 || steps[s].channel > 2)) return fail(SNES_E_INVALID, "channel out of range");
    }
    const int per_img = nsteps * ncand;
    const size_t E = (size_t)nimg * per_img;
    if (mode == 0)
        for (size_t i = 0; i < E * 3; i++)
            if (cand[i] > 32) return fail(SNES_E_INVALID, "colour component > 32");
    RET(ensure_evals(ctx, E));
    RET(ensure_ints(ctx, (size_t)per_img + 2 * nsteps + 2 * nimg));
    RET(ensure_best_m(ctx, (size_t)nimg * nsteps + nimg));   // + nimg records' worth of doubles: errors before / after (iterate)
    lay.slots = ctx->d_ints;
    lay.step_slot = lay.slots + per_img;
    lay.step_channel = lay.step_slot + nsteps;
    lay.consumed = lay.step_channel + nsteps;
    lay.chosen = lay.consumed + nimg;
    std::vector<int> h((size_t)per_img + 2 * nsteps);
    for (int s = 0; s < nsteps; s++) {
        const int slot = steps[s].palette * cfg.subpalette_size + steps[s].index;
        for (int k = 0; k < ncand; k++) h[(size_t)s * ncand + k] = slot;
        h[(size_t)per_img + s] = slot;
        h[(size_t)per_img + nsteps + s] = mode == 1 ? -1 : steps[s].channel;
    }
    cudaStream_t st = ctx->stream;
    // (pageable source: the copy is staged before the call returns, so the vector may go out of scope)
    CK(cudaMemcpyAsync(ctx->d_ints, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, st));
    if (mode == 0) CK(cudaMemcpyAsync(ctx->cand, cand, E * 3, cudaMemcpyHostToDevice, st));
    else LAUNCH(ctx, "k_make_cands", k_make_cands<<<dim3(nimg, nsteps), 64, 0, st>>>(ctx->d_imgs, 0, 0, ctx->cand, ncand, lay.step_slot, lay.step_channel));
    if (with_error && ctx->fused != 3) RET(batch_error(ctx, images, nimg));
    EvalPlan pl;
    pl.nimg = nimg;
    pl.ncand = per_img;
    pl.ovr = h[per_img];
    pl.d_slots = lay.slots;
    pl.d_cand = ctx->cand;
    pl.do_assign = pl.do_score = true;
    pl.d_scores = ctx->scores;
    pl.with_self_error = with_error && ctx->fused == 3;
    pl.no_pool = finish_later;
    pl.self_fresh = all_fresh(images, nimg);
    RET(run_plan(ctx, cfg, pl));
    if (!finish_later) LAUNCH(ctx, "k_argmin", k_argmin<<<nimg * nsteps, 128, 0, st>>>(ctx->scores, ncand, 0, ctx->best_m));
    return SNES_OK;
}

extern "C" int snes_batch_eval_candidates_multi(snes_ctx *ctx, snes_image *const *images, int nimg, const snes_step *steps, int nsteps,
                                                const uint8_t *cand, int ncand, double *scores, snes_best *best) {
    MultiLayout lay;
    RET(eval_multi(ctx, images, nimg, 0, steps, nsteps, cand, ncand, false, lay));
    cudaStream_t st = ctx->stream;
    if (scores) CK(cudaMemcpyAsync(scores, ctx->scores, sizeof(double) * nimg * nsteps * ncand, cudaMemcpyDeviceToHost, st));
    if (best) CK(cudaMemcpyAsync(best, ctx->best_m, sizeof(Best) * nimg * nsteps, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SNES_OK;
}

// `nsteps` consecutive iterations of the loop of run() (lib.rs:889-910) for every image of the batch, speculatively: all
// steps' candidates are evaluated against the current state in one launch sequence, then the steps are taken in order
// until the first one that accepts a candidate (k_apply_first_accept).  The iterations before it found nothing better, so
// the state they would have left is the one they were evaluated against: the result is the reference's, iteration by
// iteration.  The trailing optimize() / error() of each iteration (lib.rs:906-910) recompute what is already known: the
// accepted evaluation's palette_map is adopted from the scratch maps (k_adopt_map) and its score is the new error.
static int iterate_multi(snes_ctx *ctx, snes_image *const *images, int nimg, int mode, const snes_step *steps, int nsteps, const uint8_t *cand,
                         int ncand, int *consumed, double *errors_before, double *errors_after) {
    MultiLayout lay;
    // one finishing kernel (error() of the images, first minima, first accept) after the candidates' pooling when the scorer
    // hands over the images' own partial sums; else the separate kernels
    const bool fuse_finish = ctx->fused == 3;
    RET(eval_multi(ctx, images, nimg, mode, steps, nsteps, cand, ncand, mode != 1, lay, fuse_finish));
    const snes_config cfg = images[0]->cfg;
    cudaStream_t st = ctx->stream;
    const int per_img = nsteps * ncand;
    double *d_before = reinterpret_cast<double *>(ctx->best_m + (size_t)nimg * nsteps), *d_after = d_before + nimg;
    if (mode == 1) CK(cudaMemsetAsync(d_before, 0xff, sizeof(double) * nimg, st));   // NaN unless a NES step left its entry as it was
    if (fuse_finish) {
        LAUNCH(ctx, "k_pool_fused", k_pool_fused<<<(nimg * per_img + 3) / 4, 128, 0, st>>>(ctx->partials, nimg * per_img, ctx->scores));
        LAUNCH(ctx, "k_finish_iterate", k_finish_iterate<<<nimg, 128, 0, st>>>(ctx->d_imgs, mode != 1 ? ctx->self_partials : nullptr, ctx->scores,
                                                                           lay.step_slot, nsteps, ctx->cand, ncand, mode == 1, ctx->best_m,
                                                                           lay.consumed, lay.chosen, d_before, d_after));
        if (errors_before) CK(cudaMemcpyAsync(errors_before, d_before, sizeof(double) * nimg, cudaMemcpyDeviceToHost, st));
    } else {
        // error() of the state the steps were evaluated against (lib.rs:199, 294); NES steps do not compute it (lib.rs:250)
        if (errors_before && mode != 1) CK(cudaMemcpyAsync(errors_before, ctx->self_scores, sizeof(double) * nimg, cudaMemcpyDeviceToHost, st));
        LAUNCH(ctx, "k_apply_first_accept", k_apply_first_accept<<<(nimg + 127) / 128, 128, 0, st>>>(ctx->d_imgs, nimg, lay.step_slot, nsteps, ctx->cand, ncand,
                                                                               ctx->best_m, mode == 1, lay.consumed, lay.chosen,
                                                                               mode == 1 ? d_before : nullptr));
        if (errors_before && mode == 1) CK(cudaMemcpyAsync(errors_before, d_before, sizeof(double) * nimg, cudaMemcpyDeviceToHost, st));
    }
    bool fresh = cfg.subpalette_count * cfg.subpalette_size <= 255 && (size_t)nimg * per_img <= (size_t)ctx->chunk;
    for (int j = 0; j < nimg; j++) fresh = fresh && images[j]->map_fresh;
    if (fresh) {
        LAUNCH(ctx, "k_adopt_map", k_adopt_map<<<dim3(64, nimg), 256, 0, st>>>(ctx->d_imgs, lay.chosen, ctx->maps, per_img, cfg.subpalette_size));
    } else {
        RET(batch_optimize(ctx, images, nimg));   // stale palette_map, or the scratch maps of the first chunks are gone
    }
    if (consumed) CK(cudaMemcpyAsync(consumed, lay.consumed, sizeof(int) * nimg, cudaMemcpyDeviceToHost, st));
    if (errors_after && fuse_finish) {
        CK(cudaMemcpyAsync(errors_after, d_after, sizeof(double) * nimg, cudaMemcpyDeviceToHost, st));
    } else if (errors_after) {
        LAUNCH(ctx, "k_load_cur_err", k_load_cur_err<<<(nimg + 127) / 128, 128, 0, st>>>(ctx->d_imgs, nimg, ctx->self_scores));
        CK(cudaMemcpyAsync(errors_after, ctx->self_scores, sizeof(double) * nimg, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return SNES_OK;
}

extern "C" int snes_image_iterate(snes_image *im, int mode, const snes_step *steps, int nsteps, const uint8_t *cand, int ncand,
                                  int *consumed, double *error_before, double *error_after) {
    if (!im) return fail(SNES_E_INVALID, "image is NULL");
    snes_image *one[1] = {im};
    return iterate_multi(im->ctx, one, 1, mode, steps, nsteps, cand, ncand, consumed, error_before, error_after);
}

extern "C" int snes_image_optimize_palette_entry_random(snes_image *im, int palette, int index, const uint8_t *cand, int ncand) {
    if (!im) return fail(SNES_E_INVALID, "image is NULL");
    snes_image *one[1] = {im};
    return batch_step(im->ctx, one, 1, palette, index, 0, 0, cand, ncand, nullptr, nullptr);
}
extern "C" int snes_image_optimize_palette_entry_nes(snes_image *im, int palette, int index) {
    if (!im) return fail(SNES_E_INVALID, "image is NULL");
    snes_image *one[1] = {im};
    return batch_step(im->ctx, one, 1, palette, index, 1, 0, nullptr, 0, nullptr, nullptr);
}
extern "C" int snes_image_optimize_palette_entry_channel(snes_image *im, int palette, int index, int channel) {
    if (!im) return fail(SNES_E_INVALID, "image is NULL");
    snes_image *one[1] = {im};
    return batch_step(im->ctx, one, 1, palette, index, 2, channel, nullptr, 0, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------------
// colour primitives for tests
// ------------------------------------------------------------------------------------------------
__global__ void k_closest(const uint8_t *colors5, int ncolors, const double *targets, int n, int cielab, const float4 *labtab,
                          int32_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t8[3];
    for (int c = 0; c < 3; c++) {  // lib.rs:773-778
        double v = targets[3 * i + c];
        v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
        v = round(v);
        t8[c] = (v != v) ? 0 : (int)v;
    }
    int bi = 0;
    if (cielab) {
        float tl, ta, tb;
        srgb8_to_lab(t8[0], t8[1], t8[2], tl, ta, tb);
        float best = __int_as_float(0x7f800000);
        for (int j = 0; j < ncolors; j++) {
            const float4 l = labtab[bgr555_index(colors5[3 * j], colors5[3 * j + 1], colors5[3 * j + 2])];
            const float d = ciede2000(l.x, l.y, l.z, tl, ta, tb);
            if (d < best) {
                best = d;
                bi = j;
            }
        }
    } else {
        int best = 0x7fffffff;
        for (int j = 0; j < ncolors; j++) {
            const uchar4 c = snes_as_rgba(colors5[3 * j], colors5[3 * j + 1], colors5[3 * j + 2]);
            const int key = redmean_key(c.x, c.y, c.z, t8[0], t8[1], t8[2]);
            if (key < best) {
                best = key;
                bi = j;
            }
        }
    }
    out[i] = bi;
}

__global__ void k_nes_only(const uint8_t *colors5, int n, int cielab, const float4 *labtab, uint8_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t o[3];
    new_nes_only(colors5 + 3 * i, cielab, labtab, o);
    out[3 * i] = o[0];
    out[3 * i + 1] = o[1];
    out[3 * i + 2] = o[2];
}

extern "C" int snes_closest_color_index(snes_ctx *ctx, const uint8_t *colors5, int ncolors, const double *targets, int n, int cielab,
                                        int32_t *out_index) {
    if (!ctx || !colors5 || !targets || !out_index || ncolors < 1 || n < 1) return fail(SNES_E_INVALID, "snes_closest_color_index: bad argument");
    for (int i = 0; i < ncolors * 3; i++)
        if (colors5[i] > 32) return fail(SNES_E_INVALID, "colour component > 32");
    RET(set_device(ctx));
    uint8_t *d_c = nullptr;
    double *d_t = nullptr;
    int32_t *d_o = nullptr;
    cudaStream_t st = ctx->stream;
    auto body = [&]() -> int {
        CK(cudaMalloc((void **)&d_c, (size_t)ncolors * 3));
        CK(cudaMalloc((void **)&d_t, sizeof(double) * 3 * n));
        CK(cudaMalloc((void **)&d_o, sizeof(int32_t) * n));
        CK(cudaMemcpyAsync(d_c, colors5, (size_t)ncolors * 3, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_t, targets, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
        LAUNCH(ctx, "k_closest", k_closest<<<(n + 127) / 128, 128, 0, st>>>(d_c, ncolors, d_t, n, cielab, ctx->labtab, d_o));
        CK(cudaMemcpyAsync(out_index, d_o, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return SNES_OK;
    };
    const int rc = body();
    cudaFree(d_c);
    cudaFree(d_t);
    cudaFree(d_o);
    return rc;
}

extern "C" int snes_new_nes_only(snes_ctx *ctx, const uint8_t *colors5, int n, int cielab, uint8_t *out5) {
    if (!ctx || !colors5 || !out5 || n < 1) return fail(SNES_E_INVALID, "snes_new_nes_only: bad argument");
    for (int i = 0; i < n * 3; i++)
        if (colors5[i] > 32) return fail(SNES_E_INVALID, "colour component > 32");
    RET(set_device(ctx));
    uint8_t *d_c = nullptr, *d_o = nullptr;
    cudaStream_t st = ctx->stream;
    auto body = [&]() -> int {
        CK(cudaMalloc((void **)&d_c, (size_t)n * 3));
        CK(cudaMalloc((void **)&d_o, (size_t)n * 3));
        CK(cudaMemcpyAsync(d_c, colors5, (size_t)n * 3, cudaMemcpyHostToDevice, st));
        LAUNCH(ctx, "k_nes_only", k_nes_only<<<(n + 127) / 128, 128, 0, st>>>(d_c, n, cielab, ctx->labtab, d_o));
        CK(cudaMemcpyAsync(out5, d_o, (size_t)n * 3, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return SNES_OK;
    };
    const int rc = body();
    cudaFree(d_c);
    cudaFree(d_o);
    return rc;
}

extern "C" int snes_image_debug_planes(snes_image *im, float *xyb, float *mu1, float *s11) {
    if (!im) return fail(SNES_E_INVALID, "image is NULL");
    const size_t n = sizeof(float) * EVAL_XYB_FLOATS;
    if (xyb) RET(d2h(im, xyb, im->dev.xyb_rm, n));
    if (mu1) RET(d2h(im, mu1, im->dev.mu1, n));
    if (s11) RET(d2h(im, s11, im->dev.s11, n));
    return SNES_OK;
}

extern "C" int snes_image_debug_lab(snes_image *im, float *lab /* 65536*4 */) {
    if (!im || !im->dev.lab) return fail(SNES_E_INVALID, "snes_image_debug_lab: image has no Lab plane (perceptual_palettes off)");
    return d2h(im, lab, im->dev.lab, sizeof(float4) * NPIX);
}

extern "C" int snes_image_kmeans_debug(snes_image *im, double *centres /* 256*3 */, int32_t *status /* 512 */) {
    if (!im || !im->km_slab) return fail(SNES_E_INVALID, "snes_image_kmeans_debug: no k-means has run on this image");
    if (centres) RET(d2h(im, centres, im->km.centres, sizeof(double) * 3 * KM_MAXK));
    if (status) RET(d2h(im, status, im->km.status, sizeof(int) * 512));
    return SNES_OK;
}

#ifdef SNES_V2_TIMING
extern "C" int snes_debug_v2_timing(unsigned long long *out16, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, snes::g_v2_timing, sizeof(unsigned long long) * 16) != cudaSuccess) return -2;
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(snes::g_v2_timing, z, sizeof z);
    }
    return 0;
}
#endif
