// K1 — fused 3x3 max-pool NMS + candidate threshold + per-channel top-K.
//
// Replaces the reference's eager chain F.pad -> F.max_pool2d -> == -> .float() -> *
// -> torch.topk (decoder/heatmap.py:15-59) and the `score < thre_hmp` filter of
// decoder/collect.py:253, which together make ~6 full passes over the
// (N, C, H, W) heat map.  Here the map is streamed from HBM exactly once:
//
//   pass 1  nms_candidates_kernel   one warp per (plane, 128-column strip, row
//           chunk); 128-bit ld.global.nc loads, a rolling 3-row window of
//           horizontal maxima in registers, neighbours across lanes through warp
//           shuffles, zero padding at the image border (reference pads with 0,
//           not -inf).  The few survivors (peak and value >= thre) are appended
//           to a per-plane candidate list as sortable 64-bit keys.
//   pass 2  select_topk_kernel      one CTA per plane ranks the candidates by
//           (value desc, flat index asc) and writes the first K.
//
// A plane with more than kCandCap survivors (noise maps, thre <= 0, or the
// stand-alone topK_channel API) is handled inside pass 2 by an exact MSB-first
// radix selection over the plane itself, so the result is the exact top-K for
// every input — there is no approximate or host-side fallback.
#include "og_common.cuh"

#include <string.h>

namespace og {

namespace {

// tuning knobs (profiles/k1_tuning.md records the sweep on B200)
#ifndef OG_K1_ROWS
#define OG_K1_ROWS 8
#endif
#ifndef OG_K1_UNROLL
#define OG_K1_UNROLL 8
#endif
#ifndef OG_K1_THREADS
#define OG_K1_THREADS 256
#endif
constexpr int kRowsPerWarp = OG_K1_ROWS;   // rows of one warp's 128-column strip
constexpr int kUnroll = OG_K1_UNROLL;      // rows fetched ahead per lane (x 16 B in flight)
constexpr int kK1Threads = OG_K1_THREADS;
constexpr int kSelectThreads = 256;
constexpr int kRadixBins = 2048;

__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

#ifndef OG_K1_MIN_CTAS
#define OG_K1_MIN_CTAS 3
#endif
template <bool kVec4>
__global__ void __launch_bounds__(kK1Threads, OG_K1_MIN_CTAS)
nms_candidates_kernel(const float *__restrict__ heat, int planes, int H, int W, float thre,
                      uint32_t *__restrict__ cand_count, uint64_t *__restrict__ cand_keys) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int strips = (W + 127) >> 7;
    const int chunks = (H + kRowsPerWarp - 1) / kRowsPerWarp;
    const long long total = (long long)planes * chunks * strips;
    if (warp >= total) return;
    const int strip = (int)(warp % strips);
    const long long t = warp / strips;
    const int chunk = (int)(t % chunks);
    const int plane = (int)(t / chunks);

    const int x0 = (strip << 7) + (lane << 2);
    const float *__restrict__ p = heat + (size_t)plane * H * W;
    const int r_begin = chunk * kRowsPerWarp;
    const int r_end = min(H, r_begin + kRowsPerWarp);

    auto load_row = [&](int r) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);      // zero padding outside the image
        if (r >= 0 && r < H) {
            const float *row = p + (size_t)r * W;
            if (kVec4) {
                if (x0 < W) v = ldg_stream_f4(row + x0);
            } else {
                if (x0 + 0 < W) v.x = __ldg(row + x0 + 0);
                if (x0 + 1 < W) v.y = __ldg(row + x0 + 1);
                if (x0 + 2 < W) v.z = __ldg(row + x0 + 2);
                if (x0 + 3 < W) v.w = __ldg(row + x0 + 3);
            }
        }
        return v;
    };
    // streaming pass: rows r_begin .. r_end-1 lie inside the image, no bounds tests needed
    auto stream_row = [&](const float *row) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kVec4) {
            if (x0 < W) v = ldg_stream_f4(row);
        } else {
            if (x0 + 0 < W) v.x = __ldg(row + 0);
            if (x0 + 1 < W) v.y = __ldg(row + 1);
            if (x0 + 2 < W) v.z = __ldg(row + 2);
            if (x0 + 3 < W) v.w = __ldg(row + 3);
        }
        return v;
    };
    const float *row = p + (size_t)r_begin * W + x0;

    // thre > 0.  The whole 8-row chunk is held in registers (8 independent 128-bit loads in flight per
    // lane).  A row is looked at again only if some lane holds a value >= thre, and then the test
    // runs from the registers: first the two VERTICAL neighbours (this lane's own rows above and
    // below — no traffic, no shuffle), and only for a row in which some pixel survives that — a blob
    // of many rows has its vertical maxima on one or two of them — the horizontal and diagonal
    // neighbours through two shuffles of the column maxima.  The rows just outside the chunk and the
    // columns just outside the strip are fetched on demand (L2 hits: a neighbouring warp streams them).
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
        v[u] = (r_begin + u < r_end) ? stream_row(row + (size_t)u * W) : make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned hot = 0;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
        hot |= (fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)) >= thre) ? (1u << u) : 0u;
    hot = __reduce_or_sync(0xffffffffu, hot);
    if (hot == 0u) return;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 above = (hot & 1u) ? load_row(r_begin - 1) : zero4;                     // zero padding outside
    const float4 below = ((hot >> (kUnroll - 1)) & 1u) ? load_row(r_begin + kUnroll) : zero4;
    auto vertical = [&](float c, float up, float dn) {       // fmaxf semantics: a NaN neighbour is ignored
        return c >= thre && !(up > c) && !(dn > c);
    };
    auto side_column = [&](int rr, int col) {                // maximum of a column just outside the strip over 3 rows
        float m = 0.0f;                                       // zero padding
        if (col >= 0 && col < W) {
            const float a0 = rr - 1 >= 0 ? __ldg(p + (size_t)(rr - 1) * W + col) : 0.0f;
            const float a1 = __ldg(p + (size_t)rr * W + col);
            const float a2 = rr + 1 < H ? __ldg(p + (size_t)(rr + 1) * W + col) : 0.0f;
            m = max3(a0, a1, a2);
        }
        return m;
    };
    unsigned peaks = 0u;             // bit 4 u + k: pixel k of this lane's four in row u is a peak
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        if (!((hot >> u) & 1u)) continue;                    // warp-uniform
        const float4 up = u == 0 ? above : v[u == 0 ? 0 : u - 1];
        const float4 dn = u == kUnroll - 1 ? below : v[u == kUnroll - 1 ? u : u + 1];
        const float4 c = v[u];
        const bool cx = vertical(c.x, up.x, dn.x), cy = vertical(c.y, up.y, dn.y);
        const bool cz = vertical(c.z, up.z, dn.z), cw = vertical(c.w, up.w, dn.w);
        if (!__any_sync(0xffffffffu, cx || cy || cz || cw)) continue;
        const int rr = r_begin + u;
        // column maxima over the three rows; the neighbours' outer columns come by shuffle
        const float mx = max3(up.x, c.x, dn.x), my = max3(up.y, c.y, dn.y);
        const float mz = max3(up.z, c.z, dn.z), mw = max3(up.w, c.w, dn.w);
        float left = __shfl_up_sync(0xffffffffu, mw, 1);
        float right = __shfl_down_sync(0xffffffffu, mx, 1);
        if (lane == 0) left = cx ? side_column(rr, x0 - 1) : 0.0f;
        if (lane == 31) right = cw ? side_column(rr, x0 + 4) : 0.0f;
        unsigned m4 = 0u;
        if (cx && !(left > c.x) && !(my > c.x) && x0 + 0 < W) m4 |= 1u;
        if (cy && !(mx > c.y) && !(mz > c.y) && x0 + 1 < W) m4 |= 2u;
        if (cz && !(my > c.z) && !(mw > c.z) && x0 + 2 < W) m4 |= 4u;
        if (cw && !(mz > c.w) && !(right > c.w) && x0 + 3 < W) m4 |= 8u;
        peaks |= m4 << (4 * u);
    }
    if (!__any_sync(0xffffffffu, peaks != 0u)) return;
    // one atomic per warp and chunk reserves the list slots of all its peaks
    const int mine = __popc(peaks);
    int before = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, before, d);
        if (lane >= d) before += t;
    }
    const int total_peaks = __shfl_sync(0xffffffffu, before, 31);
    if (total_peaks == 0) return;
    uint32_t base = 0u;
    if (lane == 31) base = atomicAdd(&cand_count[plane], (uint32_t)total_peaks);
    base = __shfl_sync(0xffffffffu, base, 31) + (uint32_t)(before - mine);
    if (mine == 0) return;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const unsigned m4 = (peaks >> (4 * u)) & 15u;
        if (m4 == 0u) continue;
        const float vals[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((m4 >> k) & 1u) {
                if (base < (uint32_t)kCandCap)
                    cand_keys[(size_t)plane * kCandCap + base] =
                        make_key(vals[k] + 0.0f, (uint32_t)((r_begin + u) * W + x0 + k));
                ++base;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// pass 2
// ---------------------------------------------------------------------------
__device__ __forceinline__ float nms_value(const float *__restrict__ p, int H, int W, int i,
                                           bool apply_nms) {
    float v = __ldg(p + i);
    if (apply_nms) {
        const int y = i / W, x = i - y * W;
        float m = 0.0f;   // zero padding takes part in the maximum at the border only
        bool border = (y == 0) | (x == 0) | (y == H - 1) | (x == W - 1);
        if (!border) m = v;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                m = fmaxf(m, __ldg(p + yy * W + xx));
            }
        }
        v = (v == m) ? v : 0.0f;
    }
    return v + 0.0f;
}

// Rank `n` distinct keys held in shared memory and write the K smallest.
__device__ void write_ranked(const uint64_t *keys, int n, int K, float *out_score,
                             int32_t *out_index) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = keys[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (keys[j] < key) ? 1 : 0;
        if (rank < K) {
            out_score[rank] = key_value(key);
            out_index[rank] = (int32_t)(uint32_t)key;
        }
    }
    for (int r = n + threadIdx.x; r < K; r += blockDim.x) {
        out_score[r] = 0.0f;
        out_index[r] = -1;
    }
}

constexpr int kSelectPlanes = kSelectThreads / 32;     // planes per CTA of the selection kernel
constexpr int kSmallCand = 128;                        // candidates a single warp ranks

struct SelectShared {
    uint64_t keys[kCandCap];
    uint32_t hist[kRadixBins];
    uint32_t part[kSelectThreads];
    uint32_t scalar[4];
    uint64_t small[kSelectPlanes][kSmallCand];
    uint32_t n_cand[kSelectPlanes];
    uint8_t big[kSelectPlanes];
};

// One plane selected by the whole CTA: up to kCandCap listed candidates are ranked in shared
// memory (completed from the zeros when thre <= 0); a plane with more goes through the exact
// radix selection over the plane itself.  Every thread of the CTA calls this with the same arguments.
__device__ void select_plane_cta(SelectShared &sh, const float *__restrict__ heat, int H, int W, float thre, int K,
                                 int plane, uint32_t n_cand, const uint64_t *__restrict__ cand_keys,
                                 float *__restrict__ o_score, int32_t *__restrict__ o_index,
                                 int32_t *__restrict__ o_count, int apply_nms,
                                 int32_t *__restrict__ overflow_flag) {
    uint64_t *s_keys = sh.keys;
    uint32_t *s_hist = sh.hist, *s_part = sh.part, *s_scalar = sh.scalar;
    const int tid = threadIdx.x;
    if (n_cand <= (uint32_t)kCandCap) {
        const int n = (int)n_cand;
        for (int i = tid; i < n; i += blockDim.x) s_keys[i] = cand_keys[(size_t)plane * kCandCap + i];
        __syncthreads();
        if (!(thre <= 0.0f) || n >= K || heat == nullptr) {
            write_ranked(s_keys, n, K, o_score, o_index);
            if (tid == 0 && o_count) *o_count = min(n, K);
            return;
        }
        // thre <= 0 (the exact joint_dets API): pass 1 listed the POSITIVE peaks only, fewer than K.
        // Next in (value desc, index asc) order come the pixels whose NMS value is 0 — every
        // non-peak — lowest index first, so the K - n first of them complete the list; the scan
        // below stops after a few hundred pixels.  A plane that does not hold enough of them
        // (then negative values follow) goes to the radix selection.
        const uint32_t need = (uint32_t)(K - n);
        uint32_t taken = 0;
        const int lane = tid & 31, wid = tid >> 5;
        const float *__restrict__ pz = heat + (size_t)plane * H * W;
        for (int base = 0; base < H * W && taken < need; base += blockDim.x) {
            const int i = base + tid;
            const bool flag = i < H * W && nms_value(pz, H, W, i, apply_nms != 0) == 0.0f;
            const uint32_t ballot = __ballot_sync(0xffffffffu, flag);
            if (lane == 0) s_part[wid] = __popc(ballot);
            __syncthreads();
            uint32_t before = 0, chunk_total = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
                const uint32_t c = s_part[w];
                if (w < wid) before += c;
                chunk_total += c;
            }
            const uint32_t pos = taken + before + __popc(ballot & ((1u << lane) - 1u));
            if (flag && pos < need) s_keys[n + pos] = make_key(0.0f, (uint32_t)i);
            taken += chunk_total;
            __syncthreads();
        }
        if (taken >= need) {
            write_ranked(s_keys, K, K, o_score, o_index);
            if (tid == 0 && o_count) *o_count = K;
            return;
        }
    }

    if (heat == nullptr) {      // fused path: no materialised plane to re-scan; og_fetch_result
        if (tid == 0) {         // materialises the planes marked -1 and selects them exactly
            *reinterpret_cast<volatile int32_t *>(overflow_flag) = 1;      // may be mapped host memory
            if (o_count) *o_count = -1;
        }
        write_ranked(s_keys, 0, K, o_score, o_index);
        return;
    }
    // ---- exact radix selection over the plane (rare path) ----
    const float *__restrict__ p = heat + (size_t)plane * H * W;
    const int HW = H * W;
    const bool nms = apply_nms != 0;
    uint32_t prefix = 0, prefix_mask = 0;
    uint32_t remaining = (uint32_t)K;
    uint32_t total_q = 0;
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = shifts[pass];
        const uint32_t mask = (1u << widths[pass]) - 1u;
        for (int b = tid; b < kRadixBins; b += blockDim.x) s_hist[b] = 0;
        __syncthreads();
        // heat maps are mostly one value (the zero background): the lanes of a warp that hit the
        // same bin add up first, one shared-memory atomic per distinct bin and warp
        for (int base = 0; base < HW; base += blockDim.x) {
            const int i = base + tid;
            uint32_t bin = 0xffffffffu;
            if (i < HW) {
                const float nv = nms_value(p, H, W, i, nms);
                if (nv >= thre) {
                    const uint32_t key = ~ordered_bits(nv);
                    if ((key & prefix_mask) == prefix) bin = (key >> shift) & mask;
                }
            }
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (bin != 0xffffffffu && (tid & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[bin], (uint32_t)__popc(peers));
        }
        __syncthreads();
        // 256 partial sums of 8 bins each, then thread 0 walks them
        {
            uint32_t acc = 0;
            for (int b = 0; b < kRadixBins / kSelectThreads; ++b)
                acc += s_hist[tid * (kRadixBins / kSelectThreads) + b];
            s_part[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t cum = 0, tot = 0;
            for (int t = 0; t < kSelectThreads; ++t) tot += s_part[t];
            if (pass == 0) {
                s_scalar[2] = tot;
                if (tot < remaining) remaining = tot;       // fewer than K qualify
            }
            uint32_t chosen = 0;
            if (remaining > 0) {
                int t = 0;
                while (t < kSelectThreads - 1 && cum + s_part[t] < remaining) cum += s_part[t++];
                int b = t * (kRadixBins / kSelectThreads);
                while (b < kRadixBins - 1 && cum + s_hist[b] < remaining) cum += s_hist[b++];
                chosen = (uint32_t)b;
            }
            s_scalar[0] = chosen;
            s_scalar[1] = remaining - cum;     // still needed inside the chosen bin
        }
        __syncthreads();
        prefix |= s_scalar[0] << shift;
        prefix_mask |= mask << shift;
        remaining = s_scalar[1];
        total_q = s_scalar[2];
        __syncthreads();
    }
    const uint32_t k_eff = min((uint32_t)K, total_q);
    if (k_eff == 0) {
        write_ranked(s_keys, 0, K, o_score, o_index);
        if (tid == 0 && o_count) *o_count = 0;
        return;
    }
    const uint32_t kth = prefix;                 // key of the k_eff-th element
    const uint32_t n_equal_needed = remaining;   // how many of value == kth to take
    const uint32_t n_less = k_eff - n_equal_needed;

    if (tid == 0) s_scalar[3] = 0;
    __syncthreads();
    for (int i = tid; i < HW; i += blockDim.x) {
        const float nv = nms_value(p, H, W, i, nms);
        if (nv >= thre) {
            const uint32_t key = ~ordered_bits(nv);
            if (key < kth) {
                const uint32_t pos = atomicAdd(&s_scalar[3], 1u);
                s_keys[pos] = ((uint64_t)key << 32) | (uint32_t)i;
            }
        }
    }
    __syncthreads();
    // the lowest-index elements among those equal to the k-th value, in index order
    uint32_t taken = 0;
    const int lane = tid & 31, wid = tid >> 5;
    for (int base = 0; base < HW && taken < n_equal_needed; base += blockDim.x) {
        const int i = base + tid;
        bool flag = false;
        if (i < HW) {
            const float nv = nms_value(p, H, W, i, nms);
            flag = (nv >= thre) && (~ordered_bits(nv) == kth);
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_part[wid] = __popc(ballot);
        __syncthreads();
        uint32_t before = 0, chunk_total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            const uint32_t c = s_part[w];
            if (w < wid) before += c;
            chunk_total += c;
        }
        const uint32_t pos = taken + before + __popc(ballot & ((1u << lane) - 1u));
        if (flag && pos < n_equal_needed) s_keys[n_less + pos] = ((uint64_t)kth << 32) | (uint32_t)i;
        taken += chunk_total;
        __syncthreads();
    }
    write_ranked(s_keys, (int)k_eff, K, o_score, o_index);
    if (tid == 0 && o_count) *o_count = (int)k_eff;
}


// Pass 2.  A CTA takes `planes_per_cta` planes (kSelectPlanes, or 1 when long lists are expected
// everywhere: thre <= 0, forced radix selection).  First every warp ranks its own plane if that
// plane listed at most kSmallCand candidates (every real heat map: a handful per plane) — keys in
// a per-warp slice of shared memory, rank = number of smaller keys, no CTA barrier; then the CTA
// goes through the planes that need more (long lists, the zero completion of thre <= 0, the radix
// selection) one after the other.
__global__ void __launch_bounds__(kSelectThreads)
select_topk_kernel(const float *__restrict__ heat, int planes, int H, int W, float thre, int K,
                   uint32_t *__restrict__ cand_count,
                   const uint64_t *__restrict__ cand_keys, float *__restrict__ out_score,
                   int32_t *__restrict__ out_index, int32_t *__restrict__ out_count,
                   int force_radix, int apply_nms, int32_t *__restrict__ overflow_flag,
                   int32_t *__restrict__ clear_word, const int32_t *__restrict__ plane_map, int planes_per_cta) {
    __shared__ SelectShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int plane0 = blockIdx.x * planes_per_cta;
    // results of plane b of `heat` go to slot plane_map[b] of the output arrays (the redo of single
    // overflowed planes of a fused decode); the candidate lists are not used then (force_radix)
    auto out_plane = [&](int plane) { return plane_map ? plane_map[plane] : plane; };
    {
        const int plane = plane0 + warp;
        bool big = false;
        uint32_t n_cand = 0;
        if (warp < planes_per_cta && plane < planes) {
            // The plane's counter is read once and left at zero for the next call on these lists
            // (they belong to a result slot); the fused path's active-block counter likewise.
            if (lane == 0) {
                n_cand = force_radix ? (uint32_t)kCandCap + 1u : cand_count[plane];
                if (!force_radix) cand_count[plane] = 0u;
                if (clear_word != nullptr && plane == 0) *clear_word = 0;
            }
            n_cand = __shfl_sync(0xffffffffu, n_cand, 0);
            const int n = (int)n_cand;
            big = n_cand > (uint32_t)kSmallCand || (thre <= 0.0f && n < K && heat != nullptr);
            if (!big) {
                uint64_t *keys = sh.small[warp];
                for (int i = lane; i < n; i += 32) keys[i] = cand_keys[(size_t)plane * kCandCap + i];
                __syncwarp();
                const int op = out_plane(plane);
                float *o_score = out_score + (size_t)op * K;
                int32_t *o_index = out_index + (size_t)op * K;
                for (int i = lane; i < n; i += 32) {
                    const uint64_t key = keys[i];
                    int rank = 0;
                    for (int j = 0; j < n; ++j) rank += (keys[j] < key) ? 1 : 0;
                    if (rank < K) {
                        o_score[rank] = key_value(key);
                        o_index[rank] = (int32_t)(uint32_t)key;
                    }
                }
                for (int r = n + lane; r < K; r += 32) {
                    o_score[r] = 0.0f;
                    o_index[r] = -1;
                }
                if (lane == 0 && out_count) out_count[op] = min(n, K);
            }
        }
        if (lane == 0) {
            sh.big[warp] = big ? 1 : 0;
            sh.n_cand[warp] = n_cand;
        }
    }
    __syncthreads();
    for (int w = 0; w < planes_per_cta; ++w) {
        if (!sh.big[w]) continue;                          // CTA-uniform
        const int plane = plane0 + w, op = out_plane(plane);
        select_plane_cta(sh, heat, H, W, thre, K, plane, sh.n_cand[w], cand_keys, out_score + (size_t)op * K,
                         out_index + (size_t)op * K, out_count ? out_count + op : nullptr, apply_nms, overflow_flag);
        __syncthreads();
    }
}

__global__ void hmp_nms_kernel(const float *__restrict__ heat, float *__restrict__ out,
                               int planes, int H, int W) {
    const long long total = (long long)planes * H * W;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(g % W);
        const int y = (int)((g / W) % H);
        const float *p = heat + (g - (long long)y * W - x);
        const float v = p[y * W + x];
        float m = 0.0f;
        if (y > 0 && x > 0 && y < H - 1 && x < W - 1) m = v;
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                m = fmaxf(m, __ldg(p + yy * W + xx));
            }
        }
        out[g] = v * ((m == v) ? 1.0f : 0.0f);      // heat * keep_mask (heatmap.py:35)
    }
}

}  // namespace

int launch_hmp_nms(const float *heat, float *out, int planes, int h, int w, cudaStream_t s) {
    const long long total = (long long)planes * h * w;
    if (total == 0) return OG_OK;
    const int threads = 256;
    long long want = (total + threads - 1) / threads;
    const int blocks = (int)(want > 148LL * 32 ? 148LL * 32 : want);
    hmp_nms_kernel<<<blocks, threads, 0, s>>>(heat, out, planes, h, w);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

int launch_nms_topk(const float *heat, int planes, int h, int w, float thre, int k,
                    uint32_t *cand_count, uint64_t *cand_keys, float *out_score,
                    int32_t *out_index, int32_t *out_count, bool force_radix, bool apply_nms,
                    cudaStream_t s, int64_t *launches, cudaEvent_t after_pass1, const int32_t *plane_map) {
    if (planes == 0) return OG_OK;
    if (!force_radix) OG_TRY(launch_nms_candidates(heat, planes, h, w, thre, cand_count, cand_keys, s, launches));
    if (after_pass1) OG_CUDA_TRY(cudaEventRecord(after_pass1, s));
    prefer_chain_carveout<select_topk_kernel>();
    const int ppc = (force_radix || !(thre > 0.0f)) ? 1 : kSelectPlanes;
    select_topk_kernel<<<(planes + ppc - 1) / ppc, kSelectThreads, 0, s>>>(heat, planes, h, w, thre, k, cand_count, cand_keys,
                                                                         out_score, out_index, out_count,
                                                                         force_radix ? 1 : 0, apply_nms ? 1 : 0, nullptr,
                                                                         nullptr, plane_map, ppc);
    OG_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return OG_OK;
}

static inline float __int_as_float_host(int bits) {
    float f;
    memcpy(&f, &bits, sizeof(f));
    return f;
}

// pass 1 alone: clear the counters, stream the maps, append the survivors
int launch_nms_candidates(const float *heat, int planes, int h, int w, float thre,
                          uint32_t *cand_count, uint64_t *cand_keys, cudaStream_t s,
                          int64_t *launches) {
    if (planes == 0) return OG_OK;
    OG_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(uint32_t) * planes, s));
    const int strips = (w + 127) / 128;
    const int chunks = (h + kRowsPerWarp - 1) / kRowsPerWarp;
    const long long warps = (long long)planes * strips * chunks;
    const int threads = kK1Threads;
    const long long blocks = (warps + (threads / 32) - 1) / (threads / 32);
    const bool vec4 = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(heat) & 15) == 0);
    // thre <= 0 (exact joint_dets API): every pixel qualifies, but all that can rank above the
    // zeros of the non-peaks are the positive peaks — pass 1 lists those (threshold = the smallest
    // positive float) and the selection completes the list from the zeros (select_topk_kernel)
    if (!(thre > 0.0f)) thre = __int_as_float_host(1);
    auto kern = vec4 ? nms_candidates_kernel<true> : nms_candidates_kernel<false>;
    prefer_chain_carveout<nms_candidates_kernel<true>>();
    prefer_chain_carveout<nms_candidates_kernel<false>>();
    kern<<<(unsigned)blocks, threads, 0, s>>>(heat, planes, h, w, thre, cand_count, cand_keys);
    OG_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return OG_OK;
}

int launch_select_topk(const float *heat, int planes, int h, int w, float thre, int k,
                       uint32_t *cand_count, const uint64_t *cand_keys, float *out_score,
                       int32_t *out_index, int32_t *out_count, int32_t *overflow_flag,
                       int32_t *clear_word, cudaStream_t s) {
    if (planes == 0) return OG_OK;
    prefer_chain_carveout<select_topk_kernel>();
    const int ppc = !(thre > 0.0f) ? 1 : kSelectPlanes;
    select_topk_kernel<<<(planes + ppc - 1) / ppc, kSelectThreads, 0, s>>>(heat, planes, h, w, thre, k, cand_count, cand_keys,
                                                                         out_score, out_index, out_count, 0, 1,
                                                                         overflow_flag, clear_word, nullptr, ppc);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

}  // namespace og
