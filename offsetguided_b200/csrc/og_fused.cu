// K1f — flip fusion + x S resize + 3x3 NMS + threshold fused into one pass over the
// NETWORK-RESOLUTION heat maps.
//
// The reference materialises the x4 bicubic map (decoder/factory.py:74-75, 27.9 MB per
// image written, then re-read ~6 times by hmp_NMS / topk).  Here a CTA loads a 16 x 32
// low-resolution tile (+2 halo cells, fused with the mirrored copy when flip-testing) into
// shared memory, interpolates rows along x into shared memory once, and every thread
// produces its full-resolution values with one 4-tap (2-tap) combine along y.  Values are
// bit-identical to the materialised map (same taps, same accumulation order as ATen's CPU
// kernel), so the candidates are too.  A value >= thre (rare) triggers the 3x3 test, which
// recomputes the eight neighbours from the same shared rows; full-resolution maps are never
// written.  HBM traffic: C*h*w*4 bytes per image (1.74 MB instead of 55.7 MB).
#include "og_common.cuh"
#include "og_interp.cuh"

namespace og {

namespace {

constexpr int kTileW = 32;       // low-resolution cells per tile
#ifndef OG_K1F_TILE_H
#define OG_K1F_TILE_H 16
#endif
constexpr int kTileH = OG_K1F_TILE_H;
constexpr int kFusedThreads = 256;

template <int S, bool kCubic, bool kFlip>
__global__ void __launch_bounds__(kFusedThreads)
fused_nms_candidates_kernel(const float *__restrict__ hmp, const int32_t *__restrict__ kp_flip,
                            int N, int C, int h, int w, float thre,
                            uint32_t *__restrict__ cand_count, uint64_t *__restrict__ cand_keys) {
    constexpr int HALO = kCubic ? 2 : 1;
    constexpr int TAPS = kCubic ? 4 : 2;
    constexpr int LW = kTileW + 2 * HALO, LH = kTileH + 2 * HALO;
    constexpr int XW = S * kTileW + 2, YH = S * kTileH + 2;       // incl. the 1-pixel NMS ring
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
    int bid = blockIdx.x;
    const int tx = bid % tiles_x;
    bid /= tiles_x;
    const int ty = bid % tiles_y;
    const int plane = bid / tiles_y;
    const int n = plane / C, c = plane - n * C;
    const int cx0 = tx * kTileW, cy0 = ty * kTileH;
    const int W = w * S, H = h * S;

    // 1. low-resolution tile, border cells replicated (= ATen's tap clamping), fused with
    //    the mirrored copy: (orig + flip_W(flipped)[kp_flip]) / 2   (factory.py:101-106)
    const float *a = hmp + ((size_t)n * C + c) * h * w;
    const float *b = nullptr;
    if (kFlip) b = hmp + ((size_t)(N + n) * C + kp_flip[c]) * h * w;
    float tile_amax = 0.0f;
    for (int i = tid; i < LH * LW; i += kFusedThreads) {
        const int ly = i / LW, lx = i - ly * LW;
        const int gy = min(max(cy0 - HALO + ly, 0), h - 1);
        const int gx = min(max(cx0 - HALO + lx, 0), w - 1);
        float v = __ldg(a + gy * w + gx);
        if (kFlip) v = __fmul_rn(__fadd_rn(v, __ldg(b + gy * w + (w - 1 - gx))), 0.5f);
        s_lo[ly][lx] = v;
        tile_amax = fmaxf(tile_amax, fabsf(v));
    }
    // Threshold first, at network resolution: an interpolated value is bounded by
    // (sum |wx|)(sum |wy|) max|taps| <= 1.375^2 max|taps| for the A = -0.75 cubic (1 for
    // bilinear), so a tile (or, below, a cell) whose taps are all < thre / kBound cannot hold
    // a candidate and is skipped; kBound leaves 3 % for rounding.  Nothing else is skipped.
    constexpr float kBound = kCubic ? 1.95f : 1.001f;
    if (!__syncthreads_or(tile_amax * kBound >= thre)) return;

    // 2. tap tables of the full-resolution columns / rows this tile produces, and the
    //    (2 HALO + 1)^2 neighbourhood maximum of |cell| (the taps a cell's pixels can touch)
    const float inv = 1.0f / (float)S;
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
    for (int i = tid; i < LH * LW; i += kFusedThreads) {          // horizontal window maximum
        const int ly = i / LW, lx = i - ly * LW;
        float m = 0.0f;
        for (int dx = -HALO; dx <= HALO; ++dx) {
            const int xx = min(max(lx + dx, 0), LW - 1);
            m = fmaxf(m, fabsf(s_lo[ly][xx]));
        }
        s_am[ly][lx] = m;
    }
    __syncthreads();
    // 3. interpolate every tile row along x: one thread per output column, taps in registers
    for (int j = tid; j < XW; j += kFusedThreads / 2) {
        if (tid >= kFusedThreads / 2) break;
        const int xb = s_xb[j];
        const float w0 = s_xw[0][j], w1 = s_xw[1][j], w2 = s_xw[TAPS - 2][j], w3 = s_xw[TAPS - 1][j];
        for (int ly = 0; ly < LH; ++ly) {
            const float *row = &s_lo[ly][xb];
            s_hb[ly][j] = kCubic ? combine4(row[0], row[1], row[2], row[3], w0, w1, w2, w3)
                                 : combine2(row[0], row[1], w0, w1);
        }
    }
    if (tid >= kFusedThreads / 2) {                                // meanwhile: vertical window maximum
        for (int i = tid - kFusedThreads / 2; i < kTileH * kTileW; i += kFusedThreads / 2) {
            const int cy = i / kTileW, cx = i - cy * kTileW;
            float m = 0.0f;
            for (int dy = 0; dy <= 2 * HALO; ++dy) m = fmaxf(m, s_am[cy + dy][cx + HALO]);
            s_act[cy][cx] = (m * kBound >= thre) ? 1 : 0;
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
    //    the few survivors.  A warp owns whole cell rows: the S output rows of a cell row share
    //    TAPS + 1 rows of s_hb and their tap tables stay in registers while the lanes sweep the
    //    columns.
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
            yoff[p] = s_yb[jy] - yb0;            // 0 or 1: floor(src) changes at most once per cell
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
                    v = kCubic ? combine4(r[0], r[1], r[TAPS - 2], r[TAPS - 1], wv[p][0], wv[p][1],
                                          wv[p][TAPS - 2], wv[p][TAPS - 1])
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
}

template <int S>
int launch_s(const float *hmp, const int32_t *kp_flip, int n, int c, int h, int w, bool cubic,
             bool flip, float thre, uint32_t *cand_count, uint64_t *cand_keys, cudaStream_t s) {
    const int tiles = ((w + kTileW - 1) / kTileW) * ((h + kTileH - 1) / kTileH);
    const long long blocks = (long long)n * c * tiles;
    if (blocks > 0x7fffffffLL) {
        set_error("fused K1: grid of %lld blocks is too large", blocks);
        return OG_ERR_INVALID_ARGUMENT;
    }
    const dim3 grid((unsigned)blocks);
    if (cubic) {
        if (flip) fused_nms_candidates_kernel<S, true, true><<<grid, kFusedThreads, 0, s>>>(hmp, kp_flip, n, c, h, w, thre, cand_count, cand_keys);
        else fused_nms_candidates_kernel<S, true, false><<<grid, kFusedThreads, 0, s>>>(hmp, kp_flip, n, c, h, w, thre, cand_count, cand_keys);
    } else {
        if (flip) fused_nms_candidates_kernel<S, false, true><<<grid, kFusedThreads, 0, s>>>(hmp, kp_flip, n, c, h, w, thre, cand_count, cand_keys);
        else fused_nms_candidates_kernel<S, false, false><<<grid, kFusedThreads, 0, s>>>(hmp, kp_flip, n, c, h, w, thre, cand_count, cand_keys);
    }
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

}  // namespace

bool fused_scale_supported(int scale) { return scale == 2 || scale == 4 || scale == 8; }

int launch_fused_candidates(const float *hmp, const int32_t *kp_flip_dev, int n, int c, int h, int w,
                            int scale, bool cubic, bool flip, float thre, uint32_t *cand_count,
                            uint64_t *cand_keys, cudaStream_t s) {
    if (n == 0) return OG_OK;
    OG_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(uint32_t) * (size_t)n * c, s));
    switch (scale) {
        case 2: return launch_s<2>(hmp, kp_flip_dev, n, c, h, w, cubic, flip, thre, cand_count, cand_keys, s);
        case 4: return launch_s<4>(hmp, kp_flip_dev, n, c, h, w, cubic, flip, thre, cand_count, cand_keys, s);
        case 8: return launch_s<8>(hmp, kp_flip_dev, n, c, h, w, cubic, flip, thre, cand_count, cand_keys, s);
        default:
            set_error("fused K1: scale %d is not instantiated", scale);
            return OG_ERR_UNSUPPORTED;
    }
}

}  // namespace og
