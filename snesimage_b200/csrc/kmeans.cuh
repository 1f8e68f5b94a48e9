// k-means initialisation kernels: initialize_tiles (lib.rs:79-189) and recalculate_palette
// (lib.rs:330-405), with cogset 0.2.0's Kmeans restated as Lloyd's algorithm (first-k initial
// centres, strict-< nearest centre, stop when |objective - previous| < 1e-6 or after 100 updates,
// empty cluster -> NaN centre).
//
// Points are stored as f32 triples (u8 RGB and Lab<f32> are both exact in f32) and widened to f64 on
// load, so all distances and sums are f64 like the reference's.  Order of the points is the
// reference's gather order, because it decides the initial centres.
//
// Summation order: cluster sums and the objective are reduced in a fixed tree order (deterministic,
// run-to-run identical).  For RGB points the sums are sums of integers < 2^53 and are therefore the
// exact values the reference gets in its sequential order.  For the tile-mean problem (n <= 1024,
// non-integer f64 points) EXACT mode adds sequentially in index order, bit-identical to the
// reference's loop.  Only the objective of the per-pixel problems and the Lab cluster sums can differ
// from a sequential sum, by rounding in the last bits.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "lab.cuh"

namespace snes {

constexpr int KM_MAX_ITER = 100;
constexpr double KM_TOL = 1e-6;
constexpr int KM_MAXK = 256;

struct KmScratch {     // per image
    float *pts;        // [NPIX][3] gathered pixel points, subpalette-major
    int *assign;       // [NPIX]
    int *sub_off;      // [C+1] offsets of each subpalette's points
    double *means;     // [NTILES][3] tile means (initialize_tiles)
    int *tile_map;     // [NTILES] rank -> tile index
    int *nmeans;       // [1]
    double *centres;   // [KM_MAXK][3] output centres of the last problem set (per subpalette: [C][S][3])
    int *status;       // [C] 0 ok, -1 cogset assertion (2 <= k < n) fails; [C] = iterations (debug)
};

// Block-wide exclusive scan of one int per thread (blockDim.x == 1024).  Returns the exclusive
// prefix; *total receives the block total.
__device__ __forceinline__ int block_exscan_1024(int v, int *total) {
    __shared__ int warp_sums[32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();  // protect warp_sums reuse across calls
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        warp_sums[lane] = winc - w;
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    *total = s_total;
    return warp_sums[warp] + inc - v;
}

// k_gather_points: the point list of recalculate_palette (lib.rs:333-364) for every subpalette of one
// image: tiles in ascending index, within a tile x outer / y inner, opaque pixels only; RGB as-is or
// the pixel's Lab.  grid = nimg, block 1024 (thread = tile).
__global__ void __launch_bounds__(1024) k_gather_points(const ImgDev *imgs, const KmScratch *scr, int C, int lab) {
    const ImgDev im = imgs[blockIdx.x];
    const KmScratch sc = scr[blockIdx.x];
    const int tile = threadIdx.x, tx = tile & 31, ty = tile >> 5;
    const int mine = im.tile_pal[tile];
    int cnt = 0;
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++) cnt += im.rgba[(ty * 8 + y) * W + tx * 8 + x].w > 0;
    int base = 0;
    int my_off = -1;
    for (int p = 0; p < C; p++) {
        int total;
        const int ex = block_exscan_1024(mine == p ? cnt : 0, &total);
        if (mine == p) my_off = base + ex;
        if (tile == 0) sc.sub_off[p] = base;
        base += total;
    }
    if (tile == 0) sc.sub_off[C] = base;
    if (my_off < 0) return;  // tile_palettes value >= C: belongs to no subpalette
    float *o = sc.pts + 3 * (size_t)my_off;
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++) {
            const int px = (ty * 8 + y) * W + tx * 8 + x;
            const uchar4 c = im.rgba[px];
            if (c.w == 0) continue;
            if (lab) {
                const float4 l = reinterpret_cast<const float4 *>(im.lab)[px];
                o[0] = l.x;
                o[1] = l.y;
                o[2] = l.z;
            } else {
                o[0] = (float)c.x;
                o[1] = (float)c.y;
                o[2] = (float)c.z;
            }
            o += 3;
        }
}

// k_tile_means: lib.rs:89-128.  Tiles are visited tile_x outer / tile_y inner (rank = tile_x*32 +
// tile_y) although their index is tile_y*32 + tile_x; sums are f32 in x-outer / y-inner order; a tile
// whose (sum0 + sum1) + sum2 > 0 is false is skipped.  grid = nimg, block 1024 (thread = rank).
__global__ void __launch_bounds__(1024) k_tile_means(const ImgDev *imgs, const KmScratch *scr, int lab) {
    const ImgDev im = imgs[blockIdx.x];
    const KmScratch sc = scr[blockIdx.x];
    const int rank = threadIdx.x, tx = rank >> 5, ty = rank & 31;
    float sum[3] = {0.0f, 0.0f, 0.0f};
    int count = 0;
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++) {
            const int px = (ty * 8 + y) * W + tx * 8 + x;
            const uchar4 c = im.rgba[px];
            if (c.w == 0) continue;
            if (lab) {
                const float4 l = reinterpret_cast<const float4 *>(im.lab)[px];
                sum[0] += l.x;
                sum[1] += l.y;
                sum[2] += l.z;
            } else {
                sum[0] += (float)c.x;
                sum[1] += (float)c.y;
                sum[2] += (float)c.z;
            }
            count++;
        }
    const int keep = (sum[0] + sum[1]) + sum[2] > 0.0f;
    int total;
    const int pos = block_exscan_1024(keep, &total);
    if (keep) {
        sc.means[3 * pos] = (double)sum[0] / (double)count;
        sc.means[3 * pos + 1] = (double)sum[1] / (double)count;
        sc.means[3 * pos + 2] = (double)sum[2] / (double)count;
        sc.tile_map[pos] = ty * 32 + tx;
    }
    if (rank == 0) *sc.nmeans = total;
}

// Fixed-order block reduction of up to 4 doubles per thread (blockDim.x == 1024).
template <int NV>
__device__ __forceinline__ void block_reduce_1024(double v[NV], double *s_red /* [32][NV] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    __syncthreads();
    if (lane == 0)
        for (int q = 0; q < NV; q++) s_red[warp * NV + q] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; q++) {
        double t = s_red[lane * NV + q];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        v[q] = t;
    }
}

// k_kmeans: cogset Kmeans::new(points, k) for one problem per CTA.
//   TILES == true : points = the image's tile means (f64, n = *nmeans), k = C, EXACT summation order;
//                   afterwards tile_palettes[tile_map[i]] = assign[i]  (lib.rs:133-138).
//   TILES == false: problem (image, subpalette p): points = pts[sub_off[p] .. sub_off[p+1]), k = S.
//                   intpts != 0: the points are 8-bit integers (RGB mode).
// Centres go to scr.centres[(TILES ? 0 : p*k) ..][3]; status[p] = -1 where cogset would panic.
// grid = nimg * (TILES ? 1 : C), block 1024.
//
// Centre update in ONE pass over the points (round 2; it was k passes with a block reduction each):
//   * integer points: a cluster's sums are sums of integers below 2^24, so any order gives the reference's value.  They are
//     accumulated as u32 in shared memory while the assignment pass runs: the lanes of a warp that chose the same centre are
//     found with match.any, their three components summed by redux.sync, and one lane adds the warp's share atomically.
//   * tile means (f64, n <= 1024): the reference's sequential order matters, so cluster q's sum stays ONE thread's loop over the
//     points in index order, as does the objective.  (Running the k chains on a second warp next to the objective chain, all from
//     shared memory, was measured 2.3x SLOWER on B200 -- 1.68 against 0.73 ms for one picture, profiles/r2_kmeans_ab.txt -- and
//     was dropped.)
//   * Lab points (f64 sums of f32 values): still one fixed-order tree reduction per cluster; any order-free accumulation would
//     make the sums run-to-run different.
template <bool TILES>
__global__ void __launch_bounds__(1024) k_kmeans(const ImgDev *imgs, const KmScratch *scr, int C, int k, int intpts) {
    __shared__ double s_cent[KM_MAXK][3];
    __shared__ unsigned s_acc[TILES ? 1 : KM_MAXK * 4];   // per cluster: sum r, sum g, sum b, count
    __shared__ double s_red[32 * 4];
    __shared__ double s_cost[TILES ? NTILES : 1];
    __shared__ int s_asg[TILES ? NTILES : 1];
    __shared__ double s_obj;
    const int j = TILES ? blockIdx.x : blockIdx.x / C, p = TILES ? 0 : blockIdx.x % C;
    const ImgDev im = imgs[j];
    const KmScratch sc = scr[j];
    const int tid = threadIdx.x, lane = tid & 31;
    const int off = TILES ? 0 : sc.sub_off[p];
    const int n = TILES ? *sc.nmeans : sc.sub_off[p + 1] - off;
    const float *pts = sc.pts + 3 * (size_t)off;
    int *assign = sc.assign + off;
    double *centres = sc.centres + 3 * (size_t)(p * k);
    if (!(2 <= k && k < n)) {  // cogset: assert!(2 <= k && k < data.len())
        if (tid == 0) sc.status[p] = -1;
        return;
    }
    auto point = [&](int i, double &a, double &b, double &c) {
        if (TILES) {
            a = sc.means[3 * i];
            b = sc.means[3 * i + 1];
            c = sc.means[3 * i + 2];
        } else {
            a = (double)pts[3 * i];
            b = (double)pts[3 * i + 1];
            c = (double)pts[3 * i + 2];
        }
    };
    if (tid < k) point(tid, s_cent[tid][0], s_cent[tid][1], s_cent[tid][2]);
    __syncthreads();

    double objective = 0.0;
    int iter = 0;
    for (int round = 0;; round++) {
        // ---- update_assignments + objective (+ the cluster sums of integer points) ----------------------
        if (!TILES && intpts) {
            for (int q = tid; q < 4 * k; q += 1024) s_acc[q] = 0u;
            __syncthreads();
        }
        double cost_sum = 0.0;
        for (int base = 0; base < n; base += 1024) {   // the same trip count for every lane of a warp
            const int i = base + tid;
            int mi = -1;
            double a = 0.0, b = 0.0, c = 0.0;
            if (i < n) {
                point(i, a, b, c);
                mi = 0;
                double md = __longlong_as_double(0x7ff0000000000000ll);
                for (int q = 0; q < k; q++) {
                    const double d0 = a - s_cent[q][0], d1 = b - s_cent[q][1], d2 = c - s_cent[q][2];
                    const double dd = (d0 * d0 + d1 * d1) + d2 * d2;
                    if (dd < md) {
                        md = dd;
                        mi = q;
                    }
                }
                if (TILES) {
                    s_asg[i] = mi;
                    s_cost[i] = md;
                } else {
                    if (!intpts) assign[i] = mi;
                    cost_sum += md;
                }
            }
            if (!TILES && intpts) {
                const unsigned peers = __match_any_sync(0xffffffffu, mi);
                if (mi >= 0) {   // every lane of `peers` holds the same mi, so all of them are here
                    const unsigned sr = __reduce_add_sync(peers, (unsigned)a), sg = __reduce_add_sync(peers, (unsigned)b),
                                   sb = __reduce_add_sync(peers, (unsigned)c);
                    if (lane == __ffs(peers) - 1) {
                        atomicAdd(&s_acc[4 * mi], sr);
                        atomicAdd(&s_acc[4 * mi + 1], sg);
                        atomicAdd(&s_acc[4 * mi + 2], sb);
                        atomicAdd(&s_acc[4 * mi + 3], (unsigned)__popc(peers));
                    }
                }
            }
        }
        double new_obj;
        if (TILES) {
            __syncthreads();
            if (tid == 0) {
                double o = 0.0;
                for (int i = 0; i < n; i++) o += s_cost[i];
                s_obj = o;
            }
            __syncthreads();
            new_obj = s_obj;
        } else {
            double v[1] = {cost_sum};
            block_reduce_1024<1>(v, s_red);
            new_obj = v[0];
        }
        if (round > 0) {
            if (fabs(new_obj - objective) < KM_TOL) break;
            objective = new_obj;
            iter++;
            if (iter >= KM_MAX_ITER) break;
        } else {
            objective = new_obj;
        }
        // ---- update_centres ----------------------------------------------------------------------
        __syncthreads();
        if (TILES) {
            if (tid < k) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                int cnt = 0;
                for (int i = 0; i < n; i++)
                    if (s_asg[i] == tid) {
                        s0 += sc.means[3 * i];
                        s1 += sc.means[3 * i + 1];
                        s2 += sc.means[3 * i + 2];
                        cnt++;
                    }
                const double scale = 1.0 / (double)cnt;
                s_cent[tid][0] = s0 * scale;
                s_cent[tid][1] = s1 * scale;
                s_cent[tid][2] = s2 * scale;
            }
        } else if (intpts) {
            if (tid < k) {
                const double scale = 1.0 / (double)s_acc[4 * tid + 3];
                s_cent[tid][0] = (double)s_acc[4 * tid] * scale;
                s_cent[tid][1] = (double)s_acc[4 * tid + 1] * scale;
                s_cent[tid][2] = (double)s_acc[4 * tid + 2] * scale;
            }
        } else {
            for (int q = 0; q < k; q++) {
                double v[4] = {0.0, 0.0, 0.0, 0.0};
                for (int i = tid; i < n; i += 1024)
                    if (assign[i] == q) {
                        v[0] += (double)pts[3 * i];
                        v[1] += (double)pts[3 * i + 1];
                        v[2] += (double)pts[3 * i + 2];
                        v[3] += 1.0;
                    }
                block_reduce_1024<4>(v, s_red);
                if (tid == 0) {
                    const double scale = 1.0 / v[3];
                    s_cent[q][0] = v[0] * scale;
                    s_cent[q][1] = v[1] * scale;
                    s_cent[q][2] = v[2] * scale;
                }
            }
        }
        __syncthreads();
    }
    // cogset's loop: the last statement executed is an update_assignments, so (centres, assignments)
    // are consistent here.
    __syncthreads();
    if (tid < k) {
        centres[3 * tid] = s_cent[tid][0];
        centres[3 * tid + 1] = s_cent[tid][1];
        centres[3 * tid + 2] = s_cent[tid][2];
    }
    if (TILES)
        for (int i = tid; i < n; i += 1024) im.tile_pal[sc.tile_map[i]] = (uint8_t)s_asg[i];
    if (tid == 0) {
        sc.status[p] = 0;
        sc.status[C + p] = iter;
    }
}

// k_kmeans<false> for integer points when there are few problems (one picture: C problems on 148 SMs): the points of a problem are
// spread over a thread-block cluster of KM_CLUSTER CTAs.  Every CTA assigns its share and accumulates its own integer sums and its
// share of the objective; after one cluster barrier every CTA reads all partial results through distributed shared memory in rank
// order -- the same totals and the same objective in every CTA, so all of them take the same decision and compute the same centres
// (redundantly) for the next pass.  The accumulators are double-buffered, which makes that barrier the only one of an iteration.
// Integer sums are exact in any order, so the centres are those of the one-CTA kernel; the objective is the sum of the CTAs'
// fixed-order partial sums (it only decides when to stop: |change| < 1e-6).
// grid = problems x KM_CLUSTER, block 1024.
constexpr int KM_CLUSTER = 8;
__global__ void __cluster_dims__(KM_CLUSTER, 1, 1) __launch_bounds__(1024) k_kmeans_cluster(const ImgDev *imgs, const KmScratch *scr, int C, int k) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ double s_cent[KM_MAXK][3];
    __shared__ unsigned s_acc[2][KM_MAXK * 4];   // [buffer][cluster: sum r, sum g, sum b, count] of THIS CTA's points
    __shared__ double s_costpart[2];
    __shared__ unsigned s_tot[KM_MAXK * 4];
    __shared__ double s_red[32 * 4];
    __shared__ double s_obj;
    const int prob = blockIdx.x / KM_CLUSTER, j = prob / C, p = prob % C;
    const KmScratch sc = scr[j];
    const int tid = threadIdx.x, lane = tid & 31;
    const int off = sc.sub_off[p];
    const int n = sc.sub_off[p + 1] - off;
    const float *pts = sc.pts + 3 * (size_t)off;
    double *centres = sc.centres + 3 * (size_t)(p * k);
    if (!(2 <= k && k < n)) {  // cogset: assert!(2 <= k && k < data.len())   (the same in every CTA of the cluster)
        if (rank == 0 && tid == 0) sc.status[p] = -1;
        return;
    }
    if (tid < k) {
        s_cent[tid][0] = (double)pts[3 * tid];
        s_cent[tid][1] = (double)pts[3 * tid + 1];
        s_cent[tid][2] = (double)pts[3 * tid + 2];
    }
    __syncthreads();
    double objective = 0.0;
    int iter = 0, buf = 0;
    for (int round = 0;; round++, buf ^= 1) {
        for (int q = tid; q < 4 * k; q += 1024) s_acc[buf][q] = 0u;
        __syncthreads();
        double cost_sum = 0.0;
        for (int base = 0; base < n; base += 1024 * KM_CLUSTER) {
            const int i = base + rank * 1024 + tid;
            int mi = -1;
            double a = 0.0, b = 0.0, c = 0.0;
            if (i < n) {
                a = (double)pts[3 * i];
                b = (double)pts[3 * i + 1];
                c = (double)pts[3 * i + 2];
                mi = 0;
                double md = __longlong_as_double(0x7ff0000000000000ll);
                for (int q = 0; q < k; q++) {
                    const double d0 = a - s_cent[q][0], d1 = b - s_cent[q][1], d2 = c - s_cent[q][2];
                    const double dd = (d0 * d0 + d1 * d1) + d2 * d2;
                    if (dd < md) {
                        md = dd;
                        mi = q;
                    }
                }
                cost_sum += md;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, mi);
            if (mi >= 0) {
                const unsigned sr = __reduce_add_sync(peers, (unsigned)a), sg = __reduce_add_sync(peers, (unsigned)b),
                               sb = __reduce_add_sync(peers, (unsigned)c);
                if (lane == __ffs(peers) - 1) {
                    atomicAdd(&s_acc[buf][4 * mi], sr);
                    atomicAdd(&s_acc[buf][4 * mi + 1], sg);
                    atomicAdd(&s_acc[buf][4 * mi + 2], sb);
                    atomicAdd(&s_acc[buf][4 * mi + 3], (unsigned)__popc(peers));
                }
            }
        }
        double v[1] = {cost_sum};
        block_reduce_1024<1>(v, s_red);
        if (tid == 0) s_costpart[buf] = v[0];
        cluster.sync();   // every CTA's partial results of this pass are complete and visible
        if (tid < 4 * k) {
            unsigned t = 0;
            for (int r = 0; r < KM_CLUSTER; r++) t += cluster.map_shared_rank(&s_acc[buf][0], r)[tid];
            s_tot[tid] = t;
        }
        if (tid == 0) {
            double o = 0.0;
            for (int r = 0; r < KM_CLUSTER; r++) o += *cluster.map_shared_rank(&s_costpart[buf], r);
            s_obj = o;
        }
        __syncthreads();
        const double new_obj = s_obj;
        if (round > 0) {
            if (fabs(new_obj - objective) < KM_TOL) break;
            objective = new_obj;
            iter++;
            if (iter >= KM_MAX_ITER) break;
        } else {
            objective = new_obj;
        }
        if (tid < k) {
            const double scale = 1.0 / (double)s_tot[4 * tid + 3];
            s_cent[tid][0] = (double)s_tot[4 * tid] * scale;
            s_cent[tid][1] = (double)s_tot[4 * tid + 1] * scale;
            s_cent[tid][2] = (double)s_tot[4 * tid + 2] * scale;
        }
        __syncthreads();
    }
    cluster.sync();   // no CTA leaves while its shared memory may still be read
    if (rank == 0) {
        if (tid < k) {
            centres[3 * tid] = s_cent[tid][0];
            centres[3 * tid + 1] = s_cent[tid][1];
            centres[3 * tid + 2] = s_cent[tid][2];
        }
        if (tid == 0) {
            sc.status[p] = 0;
            sc.status[C + p] = iter;
        }
    }
}

// (v).round() as u8 : half away from zero, saturating, NaN -> 0
__device__ __forceinline__ uint8_t f64_round_as_u8(double v) {
    v = round(v);
    if (!(v > 0.0)) return 0;
    if (v > 255.0) return 255;
    return (uint8_t)v;
}

// SnesColor::new_nes_only (lib.rs:640-660): nearest of the 56 NES colours, strict-< first minimum.
__device__ inline void new_nes_only(const uint8_t c5[3], int cielab, const float4 *labtab, uint8_t out[3]) {
    const uchar4 color = snes_as_rgba(c5[0], c5[1], c5[2]);
    int best = 0;
    if (cielab) {
        float L, A, B;
        srgb8_to_lab(color.x, color.y, color.z, L, A, B);
        float be = __int_as_float(0x7f800000);
        for (int i = 0; i < NES_COUNT; i++) {
            const float4 l = labtab[bgr555_index(c_nes[i][0], c_nes[i][1], c_nes[i][2])];
            const float d = ciede2000(L, A, B, l.x, l.y, l.z);  // (color, nes candidate): lib.rs:648
            if (d < be) {
                be = d;
                best = i;
            }
        }
    } else {
        int be = 0x7fffffff;
        for (int i = 0; i < NES_COUNT; i++) {
            const uchar4 c = snes_as_rgba(c_nes[i][0], c_nes[i][1], c_nes[i][2]);
            const int key = redmean_key(color.x, color.y, color.z, c.x, c.y, c.z);
            if (key < be) {
                be = key;
                best = i;
            }
        }
    }
    out[0] = c_nes[best][0];
    out[1] = c_nes[best][1];
    out[2] = c_nes[best][2];
}

// centre (f64 triple) -> SnesColor: the shared tail of lib.rs:140-171 and 369-401.
__device__ inline void centre_to_color(const double v[3], int lab, int nes, const float4 *labtab, uint8_t out[3]) {
    uint8_t c5[3];
    if (lab) {
        uint8_t rgb[3];
        lab_f64_to_srgb8(v, rgb);
        c5[0] = rgb[0] / 8;
        c5[1] = rgb[1] / 8;
        c5[2] = rgb[2] / 8;
    } else {
        c5[0] = f64_round_as_u8(v[0] / 8.0);
        c5[1] = f64_round_as_u8(v[1] / 8.0);
        c5[2] = f64_round_as_u8(v[2] / 8.0);
    }
    if (nes) new_nes_only(c5, lab, labtab, out);
    else {
        out[0] = c5[0];
        out[1] = c5[1];
        out[2] = c5[2];
    }
}

// k_centres_to_palette: write the k-means centres into the image palettes.
//   tiles != 0: centre q of the tile problem seeds all S entries of subpalette q (lib.rs:181-183);
//   tiles == 0: centre (p, i) becomes entry p*S + i (lib.rs:404), only where status[p] == 0.
// grid = nimg, block = 256 (thread = palette slot).
__global__ void k_centres_to_palette(const ImgDev *imgs, const KmScratch *scr, int C, int S, int lab, int nes,
                                     const float4 *labtab, int tiles) {
    const ImgDev im = imgs[blockIdx.x];
    const KmScratch sc = scr[blockIdx.x];
    const int slot = threadIdx.x;
    if (slot >= C * S) return;
    const int p = slot / S;
    if (sc.status[tiles ? 0 : p] != 0) return;
    const double *v = sc.centres + 3 * (size_t)(tiles ? p : slot);
    uint8_t out[3];
    centre_to_color(v, lab, nes, labtab, out);
    im.palette[3 * slot] = out[0];
    im.palette[3 * slot + 1] = out[1];
    im.palette[3 * slot + 2] = out[2];
}

}  // namespace snes
