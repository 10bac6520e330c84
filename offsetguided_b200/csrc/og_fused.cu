// K1f — flip fusion + x S resize + 3x3 NMS + threshold fused into one pass over the
// NETWORK-RESOLUTION heat maps.
//
// The reference materialises the x4 bicubic map (decoder/factory.py:74-75, 27.9 MB per
// image written, then re-read ~6 times by hmp_NMS / topk).  Here the full-resolution map is
// never written; HBM traffic is C*h*w*4 bytes per image (1.74 MB instead of 55.7 MB).
//
// Threshold first, at network resolution: an interpolated value is bounded by
// (sum |wx|)(sum |wy|) max|taps| <= 1.375^2 max|taps| for the A = -0.75 cubic (1 for bilinear),
// so a region whose taps are all below thre / kBound cannot hold a candidate.  Three kernels:
//
//   tile_scan_kernel   streams the maps once (coalesced, one CTA per 16-row band of a plane)
//                      and writes max |fused value| of every 16 x 32-cell tile;
//   tile_list_kernel   a tile is ACTIVE if it or one of its 8 neighbours (they cover its 2-cell
//                      halo) can reach thre; active tiles are appended to a work list;
//   fused_nms_candidates_kernel   persistent CTAs walk the work list; the next tile's cells are
//                      prefetched into registers while the current tile is processed.  A CTA
//                      stages the tile (+halo, fused with the mirrored copy when flip-testing)
//                      in shared memory, interpolates rows along x once, and every warp produces
//                      the full-resolution values of its ACTIVE cells with one 4-tap (2-tap)
//                      combine along y.  Values are bit-identical to the materialised map (same
//                      taps, same accumulation order as ATen's CPU kernel), so the candidates
//                      are too.  A value >= thre (rare) triggers the 3x3 test, which recomputes
//                      the neighbours from the same shared rows.
// Skipping is provably lossless (kBound leaves 3 % for rounding); nothing else is skipped.
#include "og_common.cuh"
#include "og_interp.cuh"

#include <algorithm>

namespace og {

namespace {

constexpr int kTileW = 32;       // low-resolution cells per tile
#ifndef OG_K1F_TILE_H
#define OG_K1F_TILE_H 16
#endif
constexpr int kTileH = OG_K1F_TILE_H;
constexpr int kFusedThreads = 256;
constexpr int kScanThreads = 256;

__device__ __forceinline__ float bound_factor(bool cubic) { return cubic ? 1.95f : 1.001f; }

// ---- tile scan -------------------------------------------------------------
template <bool kFlip>
__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(const float *__restrict__ hmp, const int32_t *__restrict__ kp_flip, int N, int C,
                 int h, int w, float *__restrict__ tile_amax) {
    __shared__ unsigned s_max[64];
    const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;
    const int ty = blockIdx.x % tiles_y;
    const int plane = blockIdx.x / tiles_y;
    const int n = plane / C, c = plane - n * C;
    const int tid = threadIdx.x;
    for (int i = tid; i < tiles_x && i < 64; i += kScanThreads) s_max[i] = 0u;
    __syncthreads();
    const float *a = hmp + ((size_t)n * C + c) * h * w;
    const float *b = nullptr;
    if (kFlip) b = hmp + ((size_t)(N + n) * C + kp_flip[c]) * h * w;
    const int y0 = ty * kTileH, y1 = min(h, y0 + kTileH);
    const int cells = (y1 - y0) * w;
    for (int i = tid; i < cells; i += kScanThreads) {
        const int y = y0 + i / w, x = i - (i / w) * w;
        float v = __ldg(a + y * w + x);
        if (kFlip) v = __fmul_rn(__fadd_rn(v, __ldg(b + y * w + (w - 1 - x))), 0.5f);
        // |v| >= 0: the IEEE bit pattern orders like the value (NaN sorts above everything,
        // which keeps such tiles active)
        atomicMax(&s_max[min(x / kTileW, 63)], __float_as_uint(fabsf(v)));
    }
    __syncthreads();
    for (int i = tid; i < tiles_x; i += kScanThreads)
        tile_amax[((size_t)plane * tiles_y + ty) * tiles_x + i] = __uint_as_float(s_max[min(i, 63)]);
}

__global__ void tile_list_kernel(const float *__restrict__ tile_amax, int planes, int tiles_y,
                                 int tiles_x, float limit, int32_t *__restrict__ tile_list,
                                 int32_t *__restrict__ n_active) {
    const long long total = (long long)planes * tiles_y * tiles_x;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int tx = (int)(g % tiles_x);
    const int ty = (int)((g / tiles_x) % tiles_y);
    const long long plane = g / ((long long)tiles_x * tiles_y);
    const float *p = tile_amax + plane * tiles_y * tiles_x;
    bool active = false;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = ty + dy, xx = tx + dx;
            if (yy < 0 || yy >= tiles_y || xx < 0 || xx >= tiles_x) continue;
            const float m = p[yy * tiles_x + xx];
            active = active || !(m < limit);          // NaN keeps the tile active
        }
    if (active) tile_list[atomicAdd(n_active, 1)] = (int32_t)g;
}

// ---- interpolation + NMS over the active tiles ----------------------------------
template <int S, bool kCubic, bool kFlip>
__global__ void __launch_bounds__(kFusedThreads)
fused_nms_candidates_kernel(const float *__restrict__ hmp, const int32_t *__restrict__ kp_flip,
                            int N, int C, int h, int w, float thre,
                            const int32_t *__restrict__ tile_list,
                            const int32_t *__restrict__ n_active_ptr,
                            uint32_t *__restrict__ cand_count, uint64_t *__restrict__ cand_keys) {
    constexpr int HALO = kCubic ? 2 : 1;
    constexpr int TAPS = kCubic ? 4 : 2;
    constexpr int LW = kTileW + 2 * HALO, LH = kTileH + 2 * HALO;
    constexpr int XW = S * kTileW + 2, YH = S * kTileH + 2;       // incl. the 1-pixel NMS ring
    constexpr int kLoads = (LH * LW + kFusedThreads - 1) / kFusedThreads;
    __shared__ float s_lo[LH][LW + 1];
    __shared__ float s_hb[LH][XW];
    __shared__ float s_xw[TAPS][XW];
    __shared__ float s_yw[TAPS][YH];
    __shared__ int s_xb[XW];
    __shared__ int s_yb[YH];
    __shared__ float s_am[LH][LW + 1];
    __shared__ uint8_t s_act[kTileH][kTileW];

    const int tid = threadIdx.x;
    const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;
    const int W = w * S, H = h * S;
    const int n_active = *n_active_ptr;
    const float kBound = bound_factor(kCubic);
    const float inv = 1.0f / (float)S;

    // cells of a tile (+halo), border cells replicated (= ATen's tap clamping), fused with
    // the mirrored copy: (orig + flip_W(flipped)[kp_flip]) / 2   (factory.py:101-106)
    auto load_tile = [&](int tile, float (&vals)[kLoads]) {
        const int tx = tile % tiles_x;
        const int ty = (tile / tiles_x) % tiles_y;
        const int plane = tile / (tiles_x * tiles_y);
        const int n = plane / C, c = plane - n * C;
        const float *a = hmp + ((size_t)n * C + c) * h * w;
        const float *b = nullptr;
        if (kFlip) b = hmp + ((size_t)(N + n) * C + kp_flip[c]) * h * w;
#pragma unroll
        for (int u = 0; u < kLoads; ++u) {
            const int i = tid + u * kFusedThreads;
            float v = 0.0f;
            if (i < LH * LW) {
                const int ly = i / LW, lx = i - ly * LW;
                const int gy = min(max(ty * kTileH - HALO + ly, 0), h - 1);
                const int gx = min(max(tx * kTileW - HALO + lx, 0), w - 1);
                v = __ldg(a + gy * w + gx);
                if (kFlip) v = __fmul_rn(__fadd_rn(v, __ldg(b + gy * w + (w - 1 - gx))), 0.5f);
            }
            vals[u] = v;
        }
    };

    float vals[kLoads];
    int item = blockIdx.x;
    if (item < n_active) load_tile(tile_list[item], vals);
    for (; item < n_active; item += gridDim.x) {
        const int tile = tile_list[item];
        const int tx = tile % tiles_x;
        const int ty = (tile / tiles_x) % tiles_y;
        const int plane = tile / (tiles_x * tiles_y);
        const int cx0 = tx * kTileW, cy0 = ty * kTileH;

        // 1. registers -> shared; start fetching the next tile of this CTA
#pragma unroll
        for (int u = 0; u < kLoads; ++u) {
            const int i = tid + u * kFusedThreads;
            if (i < LH * LW) s_lo[i / LW][i % LW] = vals[u];
        }
        if (item + (int)gridDim.x < n_active) load_tile(tile_list[item + gridDim.x], vals);
        // 2. tap tables of the full-resolution columns / rows this tile produces
        for (int j = tid; j < XW + YH; j += kFusedThreads) {
            float wt[4];
            if (j < XW) {
                const int first = axis_first_tap(S * cx0 - 1 + j, inv, kCubic, wt);
                s_xb[j] = first - (cx0 - HALO);
#pragma unroll
                for (int t = 0; t < TAPS; ++t) s_xw[t][j] = wt[t];
            } else {
                const int jj = j - XW;
                const int first = axis_first_tap(S * cy0 - 1 + jj, inv, kCubic, wt);
                s_yb[jj] = first - (cy0 - HALO);
#pragma unroll
                for (int t = 0; t < TAPS; ++t) s_yw[t][jj] = wt[t];
            }
        }
        __syncthreads();
        //    horizontal (2 HALO + 1)-window maximum of |cell| (the taps a cell's pixels can touch)
        for (int i = tid; i < LH * LW; i += kFusedThreads) {
            const int ly = i / LW, lx = i - ly * LW;
            float m = 0.0f;
            for (int dx = -HALO; dx <= HALO; ++dx) {
                const int xx = min(max(lx + dx, 0), LW - 1);
                m = fmaxf(m, fabsf(s_lo[ly][xx]));
            }
            s_am[ly][lx] = m;
        }
        __syncthreads();
        // 3. interpolate every tile row along x: one thread per output column, taps in
        //    registers; the other half of the CTA finishes the cell activity map meanwhile
        if (tid < kFusedThreads / 2) {
            for (int j = tid; j < XW; j += kFusedThreads / 2) {
                const int xb = s_xb[j];
                const float w0 = s_xw[0][j], w1 = s_xw[1][j], w2 = s_xw[TAPS - 2][j],
                            w3 = s_xw[TAPS - 1][j];
                for (int ly = 0; ly < LH; ++ly) {
                    const float *row = &s_lo[ly][xb];
                    s_hb[ly][j] = kCubic ? combine4(row[0], row[1], row[2], row[3], w0, w1, w2, w3)
                                         : combine2(row[0], row[1], w0, w1);
                }
            }
        } else {
            for (int i = tid - kFusedThreads / 2; i < kTileH * kTileW; i += kFusedThreads / 2) {
                const int cy = i / kTileW, cx = i - cy * kTileW;
                float m = 0.0f;
                for (int dy = 0; dy <= 2 * HALO; ++dy) m = fmaxf(m, s_am[cy + dy][cx + HALO]);
                s_act[cy][cx] = !(m * kBound < thre) ? 1 : 0;
            }
        }
        __syncthreads();

        auto value_at = [&](int jy, int jx) {
            const int yb = s_yb[jy];
            return kCubic ? combine4(s_hb[yb][jx], s_hb[yb + 1][jx], s_hb[yb + TAPS - 2][jx],
                                     s_hb[yb + TAPS - 1][jx], s_yw[0][jy], s_yw[1][jy],
                                     s_yw[TAPS - 2][jy], s_yw[TAPS - 1][jy])
                          : combine2(s_hb[yb][jx], s_hb[yb + 1][jx], s_yw[0][jy], s_yw[1][jy]);
        };
        // 4. full-resolution values of the active cells, threshold first; the 3x3 test only for
        //    the few survivors.  A warp owns whole cell rows: the S output rows of a cell row
        //    share TAPS + 1 rows of s_hb and their tap tables stay in registers while the lanes
        //    sweep the columns.
        constexpr int kWarps = kFusedThreads / 32;
        constexpr int kRowsPerWarp = kTileH / kWarps;
        static_assert(kTileH % kWarps == 0, "tile rows must split evenly over the warps");
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll 1
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
            const int cy = warp * kRowsPerWarp + rr;
            if (S * (cy0 + cy) >= H) break;
            int yoff[S];
            float wv[S][TAPS];
            const int yb0 = s_yb[cy * S + 1];
#pragma unroll
            for (int p = 0; p < S; ++p) {
                const int jy = cy * S + p + 1;
                yoff[p] = s_yb[jy] - yb0;        // 0 or 1: floor(src) changes at most once per cell
#pragma unroll
                for (int t = 0; t < TAPS; ++t) wv[p][t] = s_yw[t][jy];
            }
#pragma unroll 1
            for (int jx = lane + 1; jx <= S * kTileW; jx += 32) {
                const int X = S * cx0 + jx - 1;
                const bool active = X < W && s_act[cy][(jx - 1) / S] != 0;
                if (!__any_sync(0xffffffffu, active)) continue;
                if (!active) continue;
                float r[TAPS + 1];
#pragma unroll
                for (int t = 0; t <= TAPS; ++t) r[t] = s_hb[min(yb0 + t, LH - 1)][jx];
#pragma unroll
                for (int p = 0; p < S; ++p) {
                    const int Y = S * (cy0 + cy) + p;
                    if (Y >= H) break;
                    float v;
                    if (yoff[p] != 0) {              // warp-uniform
                        v = kCubic ? combine4(r[1], r[2], r[TAPS - 1], r[TAPS], wv[p][0], wv[p][1],
                                              wv[p][TAPS - 2], wv[p][TAPS - 1])
                                   : combine2(r[1], r[2], wv[p][0], wv[p][1]);
                    } else {
                        v = kCubic ? combine4(r[0], r[1], r[TAPS - 2], r[TAPS - 1], wv[p][0],
                                              wv[p][1], wv[p][TAPS - 2], wv[p][TAPS - 1])
                                   : combine2(r[0], r[1], wv[p][0], wv[p][1]);
                    }
                    if (v >= thre) {
                        const int jy = cy * S + p + 1;
                        bool peak = true;      // zero padding outside the image: v >= thre > 0 wins
                        for (int dy = -1; dy <= 1 && peak; ++dy)
                            for (int dx = -1; dx <= 1; ++dx) {
                                if (dy == 0 && dx == 0) continue;
                                const int yn = Y + dy, xn = X + dx;
                                if (yn < 0 || yn >= H || xn < 0 || xn >= W) continue;
                                if (value_at(jy + dy, jx + dx) > v) {
                                    peak = false;
                                    break;
                                }
                            }
                        if (peak) {
                            const uint32_t pos = atomicAdd(&cand_count[plane], 1u);
                            if (pos < (uint32_t)kCandCap)
                                cand_keys[(size_t)plane * kCandCap + pos] =
                                    make_key(v + 0.0f, (uint32_t)(Y * W + X));
                        }
                    }
                }
            }
        }
        __syncthreads();          // the next tile overwrites the shared arrays
    }
}

template <int S>
int launch_main(const float *hmp, const int32_t *kp_flip, int n, int c, int h, int w, bool cubic,
                bool flip, float thre, const int32_t *tile_list, const int32_t *n_active,
                uint32_t *cand_count, uint64_t *cand_keys, int grid, cudaStream_t s) {
    if (cubic) {
        if (flip) fused_nms_candidates_kernel<S, true, true><<<grid, kFusedThreads, 0, s>>>(hmp, kp_flip, n, c, h, w, thre, tile_list, n_active, cand_count, cand_keys);
        else fused_nms_candidates_kernel<S, true, false><<<grid, kFusedThreads, 0, s>>>(hmp, kp_flip, n, c, h, w, thre, tile_list, n_active, cand_count, cand_keys);
    } else {
        if (flip) fused_nms_candidates_kernel<S, false, true><<<grid, kFusedThreads, 0, s>>>(hmp, kp_flip, n, c, h, w, thre, tile_list, n_active, cand_count, cand_keys);
        else fused_nms_candidates_kernel<S, false, false><<<grid, kFusedThreads, 0, s>>>(hmp, kp_flip, n, c, h, w, thre, tile_list, n_active, cand_count, cand_keys);
    }
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

}  // namespace

bool fused_supported(int scale, int h, int w) {
    (void)h;
    return (scale == 2 || scale == 4 || scale == 8) && (w + kTileW - 1) / kTileW <= 64;
}

size_t fused_tile_count(int n, int c, int h, int w) {
    return (size_t)n * c * ((w + kTileW - 1) / kTileW) * ((h + kTileH - 1) / kTileH);
}

int launch_fused_candidates(const float *hmp, const int32_t *kp_flip_dev, int n, int c, int h, int w,
                            int scale, bool cubic, bool flip, float thre, uint32_t *cand_count,
                            uint64_t *cand_keys, float *tile_amax, int32_t *tile_list,
                            int32_t *n_active, int sm_count, cudaStream_t s, int64_t *launches) {
    if (n == 0) return OG_OK;
    const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;
    const size_t tiles = fused_tile_count(n, c, h, w);
    if (tiles > 0x7fffffffULL || tiles_x > 64) {
        set_error("fused K1: %zu tiles (%d per row) exceed the supported range", tiles, tiles_x);
        return OG_ERR_UNSUPPORTED;
    }
    OG_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(uint32_t) * (size_t)n * c, s));
    OG_CUDA_TRY(cudaMemsetAsync(n_active, 0, sizeof(int32_t), s));
    const int bands = n * c * tiles_y;
    if (flip) tile_scan_kernel<true><<<bands, kScanThreads, 0, s>>>(hmp, kp_flip_dev, n, c, h, w, tile_amax);
    else tile_scan_kernel<false><<<bands, kScanThreads, 0, s>>>(hmp, kp_flip_dev, n, c, h, w, tile_amax);
    OG_CUDA_TRY(cudaGetLastError());
    const float limit = thre / (cubic ? 1.95f : 1.001f);
    tile_list_kernel<<<(unsigned)((tiles + 255) / 256), 256, 0, s>>>(tile_amax, n * c, tiles_y, tiles_x,
                                                                    limit, tile_list, n_active);
    OG_CUDA_TRY(cudaGetLastError());
    const int grid = (int)std::min<size_t>(tiles, (size_t)sm_count * 4);
    int st;
    switch (scale) {
        case 2: st = launch_main<2>(hmp, kp_flip_dev, n, c, h, w, cubic, flip, thre, tile_list, n_active, cand_count, cand_keys, grid, s); break;
        case 4: st = launch_main<4>(hmp, kp_flip_dev, n, c, h, w, cubic, flip, thre, tile_list, n_active, cand_count, cand_keys, grid, s); break;
        case 8: st = launch_main<8>(hmp, kp_flip_dev, n, c, h, w, cubic, flip, thre, tile_list, n_active, cand_count, cand_keys, grid, s); break;
        default:
            set_error("fused K1: scale %d is not instantiated", scale);
            return OG_ERR_UNSUPPORTED;
    }
    if (st == OG_OK && launches) *launches += 3;
    return st;
}

}  // namespace og
