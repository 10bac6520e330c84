// K3 — greedy limb-ordered keypoint grouping, one CTA per image
// (reference decoder/group.py:39-240, which runs per image in a CPU process pool).
//
// The person table lives in shared memory as three planes per (person row, joint):
//   ids   int32   global keypoint id (-1 = unset)          pose column 5
//   score float   best limb score seen for the joint       pose column 4
//   xyvs  float4  x, y, keypoint score, keypoint scale     pose columns 0..3
// Persons are addressed through an `order` list (logical position -> table row), so
// deleting merged persons is a compaction of that small list, never a row move.
//
// The numpy program is sequential only in appearance; per limb type it is
//   (1) gate + sort + dedup the K limb rows          -> rank sort, O(K^2) compares
//   (2) match every person against every kept limb   -> one thread per person
//   (3) apply "both ends known" / "one end known"    -> one thread per person; the
//       reference's fancy-index scatter semantics (row-major pair order, last pair
//       wins, right-hand sides read before the statement) reduce to "the last kept
//       limb k that matches person m", because each thread owns its person row
//   (4) merge persons sharing exactly two ids        -> one thread per person a,
//       partner = largest b; survivors read rows that are being deleted, which
//       nobody writes
//   (5) append unclaimed limbs as new persons, using the reference's column-sum
//       rule including its (-1)+(+1) cancellation quirk (group.py:166).
// The table has `smem_rows` rows; if an image needs more (never seen outside noise
// inputs) the same CTA restarts that image with the table in a global slab of
// L*K rows, which cannot overflow (every new person consumes one limb row).
#include "og_common.cuh"

namespace og {

namespace {

#ifndef OG_K3_THREADS
#define OG_K3_THREADS 256
#endif
constexpr int kGroupThreads = OG_K3_THREADS;

enum Flag { kNValid = 0, kAnyP1, kAnyP2, kAnyMerge, kOutOffset, kNKept, kNNew, kNumFlags = 8 };

struct GroupArgs {
    int C, L, K;
    SkeletonDev sk;
    float dist_max;
    int use_scale;
    double person_thre;
    int sort_dim;
    int smem_rows;
    float *slab;     // per image: xyvs float4[PMAX*C], score float[PMAX*C], ids int[PMAX*C]
    size_t slab_stride;
};

struct Layout {
    size_t xyvs, score, ids, conn, k_int, p_i16, p_u8, p_f64, warp, flags, total;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// k_int: 10 int arrays of K; p_i16: 6 int16 arrays of PMAX
__host__ __device__ inline Layout make_layout(int C, int L, int K, int smem_rows) {
    const size_t pmax = (size_t)L * K;
    Layout lo;
    size_t at = 0;
    lo.xyvs = at;  at += (size_t)smem_rows * C * 16;
    lo.score = at; at += (size_t)smem_rows * C * 4;
    lo.ids = at;   at += (size_t)smem_rows * C * 4;
    lo.conn = at;  at += 2 * align_up((size_t)K * OG_LIMB_COLS * 4, 16);     // double buffer
    lo.k_int = at; at += (size_t)10 * K * 4;
    at = align_up(at, 8);
    lo.p_f64 = at; at += pmax * 8;
    lo.p_i16 = at; at += align_up((size_t)6 * pmax * 2, 4);
    lo.p_u8 = at;  at += align_up(pmax, 4);
    lo.warp = at;  at += 32 * 4;
    lo.flags = at; at += kNumFlags * 4;
    lo.total = align_up(at, 16);
    return lo;
}

// Exclusive scan of a predicate over [0, n): pos[i] = number of true entries before
// i (written for every i < n); returns the total.  All threads must call it.
template <typename Pred>
__device__ int block_scan(int n, Pred pred, int16_t *pos, int *s_warp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int running = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const bool f = (i < n) && pred(i);
        const unsigned ballot = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[wid] = __popc(ballot);
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < nw; ++w) {
            const int c = s_warp[w];
            if (w < wid) before += c;
            tot += c;
        }
        if (i < n) pos[i] = (int16_t)(running + before + __popc(ballot & ((1u << lane) - 1u)));
        running += tot;
        __syncthreads();
    }
    return running;
}

// numpy's float32 pairwise summation for n <= 128 (loops_utils.h.src, probed).
__device__ float numpy_sum_f32(const float *a, int n) {
    if (n < 8) {
        float s = 0.0f;
        for (int i = 0; i < n; ++i) s = __fadd_rn(s, a[i]);
        return s;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float s = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) s = __fadd_rn(s, a[i]);
    return s;
}

__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float unset_to_zero(float v) { return v == -1.0f ? 0.0f : v; }

__global__ void __launch_bounds__(kGroupThreads)
group_kernel(GroupArgs a, const float *__restrict__ limbs, float *__restrict__ out_poses,
             int capacity_rows, int32_t *__restrict__ out_offset, int32_t *__restrict__ out_count,
             int32_t *__restrict__ out_total, const int32_t *__restrict__ only_flagged) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (only_flagged != nullptr && only_flagged[blockIdx.x] == 0) return;   // done by the warp kernel
    const int C = a.C, L = a.L, K = a.K;
    const int pmax = L * K;
    const Layout lo = make_layout(C, L, K, a.smem_rows);
    const int tid = threadIdx.x, T = blockDim.x;
    const int img = blockIdx.x;

    float *s_conn_buf[2];
    s_conn_buf[0] = reinterpret_cast<float *>(smem + lo.conn);
    s_conn_buf[1] = s_conn_buf[0] + align_up((size_t)K * OG_LIMB_COLS * 4, 16) / 4;
    const int lane = tid & 31, wid = tid >> 5, nwarps = T >> 5;
    int *kbase = reinterpret_cast<int *>(smem + lo.k_int);
    int *s_sorted = kbase + 1 * K;
    int *s_kept = kbase + 3 * K;
    int *s_kind1 = kbase + 4 * K;
    int *s_kind2 = kbase + 5 * K;
    float *s_kscore = reinterpret_cast<float *>(kbase + 6 * K);
    int *s_n1 = kbase + 7 * K;
    int *s_n2 = kbase + 8 * K;
    int16_t *pbase = reinterpret_cast<int16_t *>(smem + lo.p_i16);
    int16_t *s_order = pbase + 0 * pmax;
    int16_t *s_order2 = pbase + 1 * pmax;
    int16_t *s_pk1 = pbase + 2 * pmax;
    int16_t *s_pk2 = pbase + 3 * pmax;
    int16_t *s_blast = pbase + 4 * pmax;
    int16_t *s_pos = pbase + 5 * pmax;
    uint8_t *s_del = smem + lo.p_u8;
    double *s_ps = reinterpret_cast<double *>(smem + lo.p_f64);
    int *s_warp = reinterpret_cast<int *>(smem + lo.warp);
    volatile int *s_flag = reinterpret_cast<volatile int *>(smem + lo.flags);

    const float *limbs_img = limbs + (size_t)img * L * K * OG_LIMB_COLS;

    float4 *xyvs = nullptr;
    float *score = nullptr;
    int *ids = nullptr;
    int mm = 0;

    for (int attempt = 0; attempt < 2; ++attempt) {
        int pcap;
        if (attempt == 0) {
            xyvs = reinterpret_cast<float4 *>(smem + lo.xyvs);
            score = reinterpret_cast<float *>(smem + lo.score);
            ids = reinterpret_cast<int *>(smem + lo.ids);
            pcap = a.smem_rows;
        } else {
            float *base = a.slab + (size_t)img * a.slab_stride;
            xyvs = reinterpret_cast<float4 *>(base);
            score = base + (size_t)pmax * C * 4;
            ids = reinterpret_cast<int *>(base + (size_t)pmax * C * 5);
            pcap = pmax;
        }
        mm = 0;
        int nalloc = 0;
        bool overflow = false;
        __syncthreads();

        // rows of limb type li + 1 are fetched with cp.async while type li is processed
        auto fetch_rows = [&](int li_next) {
            float *dst = s_conn_buf[li_next & 1];
            const float *src = limbs_img + (size_t)li_next * K * OG_LIMB_COLS;
            for (int i = tid; i < K * OG_LIMB_COLS; i += T) {
                const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst + i);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(src + i));
            }
            asm volatile("cp.async.commit_group;");
        };
        fetch_rows(0);
        for (int li = 0; li < L && !overflow; ++li) {
            const int jf = a.sk.from[li], jt = a.sk.to[li];
            const float *s_conn = s_conn_buf[li & 1];
            asm volatile("cp.async.wait_all;" ::: "memory");        // rows of type li have landed
            if (li + 1 < L) fetch_rows(li + 1);
            if (tid < kNumFlags) s_flag[tid] = 0;
            __syncthreads();

            // gate (group.py:64-76): distance test and both endpoints strictly inside the image
            auto row_valid = [&](int k) {
                const float *r = s_conn + k * OG_LIMB_COLS;
                float lim = a.dist_max;
                if (a.use_scale) lim = (r[12] != r[12]) ? r[12] : fmaxf(a.dist_max, r[12]);
                return (r[8] < lim) && (r[0] > 0.f) && (r[4] > 0.f) && (r[3] > 0.f) && (r[1] > 0.f) &&
                       (r[10] == r[10]);
            };
            // ---- A. sort the valid rows by limb score desc, ties by row asc (group.py:232,
            //         canonical stable): one warp per row, lanes over the other rows, ballots
            for (int k = wid; k < K; k += nwarps) {
                if (!row_valid(k)) continue;                       // warp-uniform
                const float sc = s_conn[k * OG_LIMB_COLS + 10];
                int rank = 0;
                for (int j0 = 0; j0 < K; j0 += 32) {
                    const int j = j0 + lane;
                    bool before = false;
                    if (j < K && row_valid(j)) {
                        const float sj = s_conn[j * OG_LIMB_COLS + 10];
                        before = sj > sc || (sj == sc && j < k);
                    }
                    rank += __popc(__ballot_sync(0xffffffffu, before));
                }
                if (lane == 0) {
                    s_sorted[rank] = k;
                    atomicAdd(const_cast<int *>(&s_flag[kNValid]), 1);
                }
            }
            __syncthreads();
            // ---- B. warp 0: keep the best row per to-joint id (group.py:233-239), compact,
            //         stage the kept rows' ids / scores
            if (wid == 0) {
                const int nvalid = s_flag[kNValid];
                int kept = 0;
                for (int r0 = 0; r0 < nvalid; r0 += 32) {
                    const int r = r0 + lane;
                    const int k = r < nvalid ? s_sorted[r] : 0;
                    const int t = r < nvalid ? (int)s_conn[k * OG_LIMB_COLS + 7] : -2;
                    bool dup = false;
                    for (int j = 0; j < kept && !dup; ++j) dup = (s_kind2[j] == t);     // earlier chunks
                    for (int q = 0; q < 31; ++q) {                                      // this chunk
                        const int tq = __shfl_sync(0xffffffffu, t, q);
                        dup = dup || (q < lane && tq == t);
                    }
                    const bool keep = r < nvalid && !dup;
                    const unsigned mask = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const int j = kept + __popc(mask & ((1u << lane) - 1u));
                        s_kept[j] = k;
                        s_kind1[j] = (int)s_conn[k * OG_LIMB_COLS + 6];
                        s_kind2[j] = t;
                        s_kscore[j] = s_conn[k * OG_LIMB_COLS + 10];
                        s_n1[j] = 0;
                        s_n2[j] = 0;
                    }
                    kept += __popc(mask);
                    __syncwarp();
                }
                if (lane == 0) s_flag[kNKept] = kept;
            }
            __syncthreads();
            const int kk = s_flag[kNKept];
            if (kk == 0) continue;                                         // group.py:84-85

            // ---- C. match persons x kept limbs on the pre-update snapshot (group.py:87-109)
            for (int m = tid; m < mm; m += T) {
                const int row = s_order[m];
                const int idf = ids[row * C + jf], idt = ids[row * C + jt];
                const float sf = score[row * C + jf], st = score[row * C + jt];
                int k1 = -1, k2 = -1;
                for (int j = 0; j < kk; ++j) {
                    const int ms = (idf == s_kind1[j] ? 1 : 0) + (idt == s_kind2[j] ? 1 : 0);
                    if (ms == 0) continue;
                    const float sc = s_kscore[j];
                    const bool rep = (sc > st) || (sc > sf);
                    if (ms == 2) {
                        atomicAdd(&s_n2[j], 1);
                        if (rep) k2 = j;
                    } else {
                        atomicAdd(&s_n1[j], 1);
                        if (rep) k1 = j;
                    }
                }
                s_pk1[m] = (int16_t)k1;
                s_pk2[m] = (int16_t)k2;
                s_blast[m] = -1;
                s_del[m] = 0;
                if (k1 >= 0) s_flag[kAnyP1] = 1;
                if (k2 >= 0) s_flag[kAnyP2] = 1;
            }
            __syncthreads();
            // ---- D. apply: both ends known (group.py:114-119), then one end known (:124-135)
            for (int m = tid; m < mm; m += T) {
                const int row = s_order[m];
                const int k2 = s_pk2[m];
                if (k2 >= 0) {
                    const float sc = s_kscore[k2];
                    score[row * C + jf] = fmaxf(sc, score[row * C + jf]);
                    score[row * C + jt] = fmaxf(sc, score[row * C + jt]);
                }
                const int k1 = s_pk1[m];
                if (k1 >= 0) {
                    const float *r = s_conn + s_kept[k1] * OG_LIMB_COLS;
                    ids[row * C + jf] = s_kind1[k1];
                    ids[row * C + jt] = s_kind2[k1];
                    xyvs[row * C + jf] = make_float4(r[0], r[1], r[2], r[11]);
                    xyvs[row * C + jt] = make_float4(r[3], r[4], r[5], r[12]);
                    const float sc = s_kscore[k1];
                    score[row * C + jf] = fmaxf(sc, score[row * C + jf]);
                    score[row * C + jt] = fmaxf(sc, score[row * C + jt]);
                }
            }
            __syncthreads();
            // ---- E. merge persons sharing exactly two keypoint ids (group.py:140-155)
            int mm_after = mm;
            if (mm >= 2) {
                for (int p = wid; p < mm; p += nwarps) {      // one warp per person, lanes over joints
                    const int rowa = s_order[p];
                    int ida[(OG_MAX_KEYPOINTS + 31) / 32];
#pragma unroll
                    for (int cc = 0; cc < (OG_MAX_KEYPOINTS + 31) / 32; ++cc) {
                        const int c = cc * 32 + lane;
                        ida[cc] = c < C ? ids[rowa * C + c] : -1;
                    }
                    for (int q = p + 1; q < mm; ++q) {
                        const int rowb = s_order[q];
                        int cnt = 0;
#pragma unroll
                        for (int cc = 0; cc < (OG_MAX_KEYPOINTS + 31) / 32; ++cc) {
                            const int c = cc * 32 + lane;
                            if (cc * 32 >= C) break;
                            const bool same = c < C && ida[cc] != -1 && ida[cc] == ids[rowb * C + c];
                            cnt += __popc(__ballot_sync(0xffffffffu, same));
                        }
                        if (cnt == 2 && lane == 0) {
                            s_blast[p] = (int16_t)q;     // ascending q: the last partner wins
                            s_del[q] = 1;
                            s_flag[kAnyMerge] = 1;
                        }
                    }
                }
                __syncthreads();
                if (s_flag[kAnyMerge]) {
                    for (int p = tid; p < mm; p += T) {
                        const int q = s_blast[p];
                        if (q < 0 || s_del[p]) continue;      // deleted rows are never read again
                        const int rowa = s_order[p], rowb = s_order[q];
                        for (int c = 0; c < C; ++c) {
                            ids[rowa * C + c] = max(ids[rowa * C + c], ids[rowb * C + c]);
                            score[rowa * C + c] = fmaxf(score[rowa * C + c], score[rowb * C + c]);
                            xyvs[rowa * C + c] = max4(xyvs[rowa * C + c], xyvs[rowb * C + c]);
                        }
                    }
                    mm_after = block_scan(mm, [&](int p) { return s_del[p] == 0; }, s_pos, s_warp);
                    for (int p = tid; p < mm; p += T)
                        if (!s_del[p]) s_order2[s_pos[p]] = s_order[p];
                    __syncthreads();
                    int16_t *tmp = s_order;
                    s_order = s_order2;
                    s_order2 = tmp;
                }
            }
            // ---- F. warp 0: unclaimed limbs start new persons (group.py:166-177), column-sum
            //         rule with its (-1) + (+1) cancellation
            if (wid == 0) {
                const int w2 = s_flag[kAnyP2] ? -1 : 2;
                const int w1 = s_flag[kAnyP1] ? -1 : 1;
                int nnew = 0;
                for (int j0 = 0; j0 < kk; j0 += 32) {
                    const int j = j0 + lane;
                    const bool isnew = j < kk && (s_n2[j] * w2 + s_n1[j] * w1 == 0);
                    const unsigned mask = __ballot_sync(0xffffffffu, isnew);
                    if (isnew) s_sorted[nnew + __popc(mask & ((1u << lane) - 1u))] = j;   // new list
                    nnew += __popc(mask);
                }
                if (lane == 0) s_flag[kNNew] = nnew;
            }
            __syncthreads();
            const int nnew = s_flag[kNNew];
            if (nalloc + nnew > pcap) {
                overflow = true;          // uniform: restart this image on the global slab
                continue;
            }
            for (int e = tid; e < nnew * C; e += T) {
                const int q = e / C, c = e - q * C;
                const int j = s_sorted[q];
                const int at = (nalloc + q) * C + c;
                const float *r = s_conn + s_kept[j] * OG_LIMB_COLS;
                if (c == 0) s_order[mm_after + q] = (int16_t)(nalloc + q);
                if (c == jt) {
                    ids[at] = s_kind2[j];
                    xyvs[at] = make_float4(r[3], r[4], r[5], r[12]);
                    score[at] = s_kscore[j];
                } else if (c == jf) {
                    ids[at] = s_kind1[j];
                    xyvs[at] = make_float4(r[0], r[1], r[2], r[11]);
                    score[at] = s_kscore[j];
                } else {
                    ids[at] = -1;
                    xyvs[at] = make_float4(-1.f, -1.f, -1.f, -1.f);
                    score[at] = -1.0f;
                }
            }
            nalloc += nnew;
            mm = mm_after + nnew;
            __syncthreads();
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (!overflow) break;
    }

    // ---- person score, threshold, stable descending sort (group.py:188-219)
    for (int m = tid; m < mm; m += T) {
        const int row = s_order[m];
        float vals[OG_MAX_KEYPOINTS];
        int n = 0;
        for (int c = 0; c < C; ++c) {
            float v;
            const float4 q = xyvs[row * C + c];
            switch (a.sort_dim) {
                case 0: v = q.x; break;
                case 1: v = q.y; break;
                case 2: v = q.z; break;
                case 3: v = q.w; break;
                case 4: v = score[row * C + c]; break;
                default: v = (float)ids[row * C + c]; break;
            }
            if (v > 0.0f) vals[n++] = v;
        }
        const double ps = (double)numpy_sum_f32(vals, n) / (double)n;      // 0/0 -> NaN, kept
        s_ps[m] = ps;
        s_del[m] = (ps < a.person_thre) ? 1 : 0;
    }
    __syncthreads();
    const int nk = block_scan(mm, [&](int m) { return s_del[m] == 0; }, s_pos, s_warp);
    for (int m = tid; m < mm; m += T)
        if (!s_del[m]) s_order2[s_pos[m]] = (int16_t)m;        // kept persons, original order
    __syncthreads();
    for (int q = tid; q < nk; q += T) {
        double ps = s_ps[s_order2[q]];
        if (ps != ps) ps = -1.0e300;        // documented deviation: NaN scores sort last
        int rank = 0;
        for (int q2 = 0; q2 < nk; ++q2) {
            double p2 = s_ps[s_order2[q2]];
            if (p2 != p2) p2 = -1.0e300;
            rank += (p2 > ps || (p2 == ps && q2 < q)) ? 1 : 0;
        }
        s_blast[q] = (int16_t)rank;
    }
    if (tid == 0) {
        const int off = atomicAdd(out_total, nk);
        s_flag[kOutOffset] = off;
        out_offset[img] = off;
        out_count[img] = nk;
    }
    __syncthreads();
    const int off = s_flag[kOutOffset];
    for (int e = tid; e < nk * C; e += T) {
        const int q = e / C, c = e - q * C;
        const int dst = off + s_blast[q];
        if (dst >= capacity_rows) continue;
        const int row = s_order[s_order2[q]];
        const float4 v = xyvs[row * C + c];
        float *o = out_poses + ((size_t)dst * C + c) * OG_POSE_COLS;
        o[0] = unset_to_zero(v.x);
        o[1] = unset_to_zero(v.y);
        o[2] = unset_to_zero(v.z);
        o[3] = unset_to_zero(v.w);
        o[4] = unset_to_zero(score[row * C + c]);
        o[5] = unset_to_zero((float)ids[row * C + c]);
    }
}


// ---------------------------------------------------------------------------
// K3w — the same algorithm, ONE WARP per image, no block barriers.
//
// Typical images have a handful of persons and K <= 64 limb rows per type, so the CTA
// kernel above spends its time waiting at barriers for single-warp phases.  Here every
// phase is warp-synchronous (shuffles, ballots, __syncwarp); each warp has an SM to itself
// (one 32-thread CTA per image).  An image that needs more than `rows` person rows raises
// its `needs_cta` flag and is re-done from scratch by group_kernel; results are identical.
// ---------------------------------------------------------------------------
constexpr int kWarpRowsMax = 64;

struct WarpLayout {
    size_t xyvs, score, ids, conn, k_int, order, ps, total;
};

__host__ __device__ inline WarpLayout make_warp_layout(int C, int K, int rows) {
    WarpLayout lo;
    size_t at = 0;
    lo.xyvs = at;  at += (size_t)rows * C * 16;
    lo.score = at; at += (size_t)rows * C * 4;
    lo.ids = at;   at += (size_t)rows * C * 4;
    lo.conn = at;  at += 2 * align_up((size_t)K * OG_LIMB_COLS * 4, 16);
    lo.k_int = at; at += (size_t)8 * K * 4;          // sorted, kept, kind1, kind2, kscore, n1, n2, newlist
    at = align_up(at, 8);
    lo.ps = at;    at += (size_t)rows * 8;
    lo.order = at; at += align_up((size_t)5 * rows * 2, 4);   // order, order2, blast, del, rank (int16)
    lo.total = align_up(at, 16);
    return lo;
}

__global__ void __launch_bounds__(32)
group_warp_kernel(GroupArgs a, int rows, const float *__restrict__ limbs,
                  float *__restrict__ out_poses, int capacity_rows,
                  int32_t *__restrict__ out_offset, int32_t *__restrict__ out_count,
                  int32_t *__restrict__ out_total, int32_t *__restrict__ needs_cta) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kMaxChunks = OG_MAX_TOPK / 32;
    const int C = a.C, L = a.L, K = a.K;
    const int lane = threadIdx.x;
    const int img = blockIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const WarpLayout lo = make_warp_layout(C, K, rows);

    float4 *xyvs = reinterpret_cast<float4 *>(smem + lo.xyvs);
    float *score = reinterpret_cast<float *>(smem + lo.score);
    int *ids = reinterpret_cast<int *>(smem + lo.ids);
    float *conn_buf[2];
    conn_buf[0] = reinterpret_cast<float *>(smem + lo.conn);
    conn_buf[1] = conn_buf[0] + align_up((size_t)K * OG_LIMB_COLS * 4, 16) / 4;
    int *kbase = reinterpret_cast<int *>(smem + lo.k_int);
    int *s_sorted = kbase + 0 * K;
    int *s_kept = kbase + 1 * K;
    int *s_kind1 = kbase + 2 * K;
    int *s_kind2 = kbase + 3 * K;
    float *s_kscore = reinterpret_cast<float *>(kbase + 4 * K);
    int *s_n1 = kbase + 5 * K;
    int *s_n2 = kbase + 6 * K;
    int *s_new = kbase + 7 * K;
    double *s_ps = reinterpret_cast<double *>(smem + lo.ps);
    int16_t *obase = reinterpret_cast<int16_t *>(smem + lo.order);
    int16_t *s_order = obase + 0 * rows;
    int16_t *s_order2 = obase + 1 * rows;
    int16_t *s_blast = obase + 2 * rows;
    int16_t *s_del = obase + 3 * rows;
    int16_t *s_rank = obase + 4 * rows;

    const float *limbs_img = limbs + (size_t)img * L * K * OG_LIMB_COLS;
    auto fetch_rows = [&](int li_next) {
        float *dst = conn_buf[li_next & 1];
        const float *src = limbs_img + (size_t)li_next * K * OG_LIMB_COLS;
        for (int i = lane; i < K * OG_LIMB_COLS; i += 32) {
            const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst + i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(src + i));
        }
        asm volatile("cp.async.commit_group;");
    };

    int mm = 0, nalloc = 0;
    bool give_up = false;
    const int kchunks = (K + 31) >> 5;
    fetch_rows(0);
    for (int li = 0; li < L; ++li) {
        const int jf = a.sk.from[li], jt = a.sk.to[li];
        const float *conn = conn_buf[li & 1];
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (li + 1 < L) fetch_rows(li + 1);

        // ---- gate (group.py:64-76) and rank sort (score desc, row asc) with shuffles
        bool valid[kMaxChunks];
        float sc[kMaxChunks];
        unsigned vmask[kMaxChunks];
        int nvalid = 0;
#pragma unroll
        for (int ch = 0; ch < kMaxChunks; ++ch) {
            valid[ch] = false;
            sc[ch] = 0.0f;
            vmask[ch] = 0;
            if (ch < kchunks) {
                const int k = ch * 32 + lane;
                if (k < K) {
                    const float *r = conn + k * OG_LIMB_COLS;
                    float lim = a.dist_max;
                    if (a.use_scale) lim = (r[12] != r[12]) ? r[12] : fmaxf(a.dist_max, r[12]);
                    valid[ch] = (r[8] < lim) && (r[0] > 0.f) && (r[4] > 0.f) && (r[3] > 0.f) &&
                                (r[1] > 0.f) && (r[10] == r[10]);
                    sc[ch] = r[10];
                }
                vmask[ch] = __ballot_sync(kFull, valid[ch]);
                nvalid += __popc(vmask[ch]);
            }
        }
        if (nvalid == 0) continue;                                  // kk == 0 (group.py:84-85)
#pragma unroll
        for (int ca = 0; ca < kMaxChunks; ++ca) {
            if (ca >= kchunks) break;
            int rank = 0;
            const int k = ca * 32 + lane;
#pragma unroll
            for (int cb = 0; cb < kMaxChunks; ++cb) {
                if (cb >= kchunks) break;
                for (int q = 0; q < 32; ++q) {
                    const float sq = __shfl_sync(kFull, sc[cb], q);
                    const int kq = cb * 32 + q;
                    const bool vq = (vmask[cb] >> q) & 1u;
                    rank += (vq && (sq > sc[ca] || (sq == sc[ca] && kq < k))) ? 1 : 0;
                }
            }
            if (valid[ca]) s_sorted[rank] = k;
        }
        __syncwarp();
        // ---- best row per to-joint id (group.py:233-239), compaction, staging
        int kk = 0;
        for (int r0 = 0; r0 < nvalid; r0 += 32) {
            const int r = r0 + lane;
            const int k = r < nvalid ? s_sorted[r] : 0;
            const int t = r < nvalid ? (int)conn[k * OG_LIMB_COLS + 7] : -2;
            bool dup = false;
            for (int j = 0; j < kk && !dup; ++j) dup = (s_kind2[j] == t);
            for (int q = 0; q < 31; ++q) {
                const int tq = __shfl_sync(kFull, t, q);
                dup = dup || (q < lane && tq == t);
            }
            const bool keep = r < nvalid && !dup;
            const unsigned mask = __ballot_sync(kFull, keep);
            if (keep) {
                const int j = kk + __popc(mask & lt_mask);
                s_kept[j] = k;
                s_kind1[j] = (int)conn[k * OG_LIMB_COLS + 6];
                s_kind2[j] = t;
                s_kscore[j] = conn[k * OG_LIMB_COLS + 10];
                s_n1[j] = 0;
                s_n2[j] = 0;
            }
            kk += __popc(mask);
            __syncwarp();
        }
        // ---- match on the pre-update snapshot (group.py:87-109) and apply (:114-135); a lane
        //      owns a person, whose updates depend on that person's own row only
        bool any_p1 = false, any_p2 = false;
        for (int m0 = 0; m0 < mm; m0 += 32) {
            const int m = m0 + lane;
            const bool live = m < mm;
            const int row = live ? s_order[m] : 0;
            const int idf = live ? ids[row * C + jf] : -3, idt = live ? ids[row * C + jt] : -3;
            const float sf = live ? score[row * C + jf] : 0.f, st = live ? score[row * C + jt] : 0.f;
            int k1 = -1, k2 = -1;
            for (int j = 0; j < kk; ++j) {
                const int ms = (idf == s_kind1[j] ? 1 : 0) + (idt == s_kind2[j] ? 1 : 0);
                const unsigned b1 = __ballot_sync(kFull, live && ms == 1);
                const unsigned b2 = __ballot_sync(kFull, live && ms == 2);
                if ((b1 | b2) == 0) continue;                       // warp-uniform
                if (lane == 0) {
                    s_n1[j] += __popc(b1);
                    s_n2[j] += __popc(b2);
                }
                if (live && ms) {
                    const float scj = s_kscore[j];
                    const bool rep = (scj > st) || (scj > sf);
                    if (rep) {
                        if (ms == 2) k2 = j; else k1 = j;
                    }
                }
            }
            any_p1 = any_p1 || __any_sync(kFull, k1 >= 0);
            any_p2 = any_p2 || __any_sync(kFull, k2 >= 0);
            if (k2 >= 0) {
                const float scj = s_kscore[k2];
                score[row * C + jf] = fmaxf(scj, score[row * C + jf]);
                score[row * C + jt] = fmaxf(scj, score[row * C + jt]);
            }
            if (k1 >= 0) {
                const float *r = conn + s_kept[k1] * OG_LIMB_COLS;
                ids[row * C + jf] = s_kind1[k1];
                ids[row * C + jt] = s_kind2[k1];
                xyvs[row * C + jf] = make_float4(r[0], r[1], r[2], r[11]);
                xyvs[row * C + jt] = make_float4(r[3], r[4], r[5], r[12]);
                const float scj = s_kscore[k1];
                score[row * C + jf] = fmaxf(scj, score[row * C + jf]);
                score[row * C + jt] = fmaxf(scj, score[row * C + jt]);
            }
        }
        __syncwarp();
        // ---- merge persons sharing exactly two ids (group.py:140-155): lanes over joints
        int mm_after = mm;
        if (mm >= 2) {
            bool any_merge = false;
            for (int p = lane; p < mm; p += 32) {
                s_blast[p] = -1;
                s_del[p] = 0;
            }
            __syncwarp();
            for (int p = 0; p < mm - 1; ++p) {
                const int rowp = s_order[p];
                int idp[(OG_MAX_KEYPOINTS + 31) / 32];
#pragma unroll
                for (int cc = 0; cc < (OG_MAX_KEYPOINTS + 31) / 32; ++cc) {
                    const int c = cc * 32 + lane;
                    idp[cc] = c < C ? ids[rowp * C + c] : -1;
                }
                for (int q = p + 1; q < mm; ++q) {
                    const int rowq = s_order[q];
                    int cnt = 0;
#pragma unroll
                    for (int cc = 0; cc < (OG_MAX_KEYPOINTS + 31) / 32; ++cc) {
                        if (cc * 32 >= C) break;
                        const int c = cc * 32 + lane;
                        const bool same = c < C && idp[cc] != -1 && idp[cc] == ids[rowq * C + c];
                        cnt += __popc(__ballot_sync(kFull, same));
                    }
                    if (cnt == 2) {
                        any_merge = true;
                        if (lane == 0) {
                            s_blast[p] = (int16_t)q;        // ascending q: the last partner wins
                            s_del[q] = 1;
                        }
                    }
                }
            }
            if (any_merge) {                                         // warp-uniform
                __syncwarp();
                for (int p = lane; p < mm; p += 32) {
                    const int q = s_blast[p];
                    if (q < 0 || s_del[p]) continue;
                    const int rowa = s_order[p], rowb = s_order[q];
                    for (int c = 0; c < C; ++c) {
                        ids[rowa * C + c] = max(ids[rowa * C + c], ids[rowb * C + c]);
                        score[rowa * C + c] = fmaxf(score[rowa * C + c], score[rowb * C + c]);
                        xyvs[rowa * C + c] = max4(xyvs[rowa * C + c], xyvs[rowb * C + c]);
                    }
                }
                int kept_persons = 0;
                for (int p0 = 0; p0 < mm; p0 += 32) {
                    const int p = p0 + lane;
                    const bool keep = p < mm && !s_del[p];
                    const unsigned mask = __ballot_sync(kFull, keep);
                    if (keep) s_order2[kept_persons + __popc(mask & lt_mask)] = s_order[p];
                    kept_persons += __popc(mask);
                }
                mm_after = kept_persons;
                int16_t *tmp = s_order;
                s_order = s_order2;
                s_order2 = tmp;
                __syncwarp();
            }
        }
        // ---- unclaimed limbs start new persons (group.py:166-177), column-sum rule
        const int w2 = any_p2 ? -1 : 2, w1 = any_p1 ? -1 : 1;
        int nnew = 0;
        for (int j0 = 0; j0 < kk; j0 += 32) {
            const int j = j0 + lane;
            const bool isnew = j < kk && (s_n2[j] * w2 + s_n1[j] * w1 == 0);
            const unsigned mask = __ballot_sync(kFull, isnew);
            if (isnew) s_new[nnew + __popc(mask & lt_mask)] = j;
            nnew += __popc(mask);
        }
        if (nalloc + nnew > rows) {          // warp-uniform: hand the image to the CTA kernel
            give_up = true;
            break;
        }
        __syncwarp();
        for (int e = lane; e < nnew * C; e += 32) {
            const int q = e / C, c = e - q * C;
            const int j = s_new[q];
            const int at = (nalloc + q) * C + c;
            const float *r = conn + s_kept[j] * OG_LIMB_COLS;
            if (c == 0) s_order[mm_after + q] = (int16_t)(nalloc + q);
            if (c == jt) {
                ids[at] = s_kind2[j];
                xyvs[at] = make_float4(r[3], r[4], r[5], r[12]);
                score[at] = s_kscore[j];
            } else if (c == jf) {
                ids[at] = s_kind1[j];
                xyvs[at] = make_float4(r[0], r[1], r[2], r[11]);
                score[at] = s_kscore[j];
            } else {
                ids[at] = -1;
                xyvs[at] = make_float4(-1.f, -1.f, -1.f, -1.f);
                score[at] = -1.0f;
            }
        }
        nalloc += nnew;
        mm = mm_after + nnew;
        __syncwarp();
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (lane == 0) needs_cta[img] = give_up ? 1 : 0;
    if (give_up) return;

    // ---- person score, threshold, stable descending sort (group.py:188-219)
    int nk = 0;
    for (int m0 = 0; m0 < mm; m0 += 32) {
        const int m = m0 + lane;
        bool keep = false;
        if (m < mm) {
            const int row = s_order[m];
            float vals[OG_MAX_KEYPOINTS];
            int n = 0;
            for (int c = 0; c < C; ++c) {
                float v;
                const float4 q = xyvs[row * C + c];
                switch (a.sort_dim) {
                    case 0: v = q.x; break;
                    case 1: v = q.y; break;
                    case 2: v = q.z; break;
                    case 3: v = q.w; break;
                    case 4: v = score[row * C + c]; break;
                    default: v = (float)ids[row * C + c]; break;
                }
                if (v > 0.0f) vals[n++] = v;
            }
            const double ps = (double)numpy_sum_f32(vals, n) / (double)n;
            s_ps[m] = ps;
            keep = !(ps < a.person_thre);
        }
        const unsigned mask = __ballot_sync(kFull, keep);
        if (keep) s_order2[nk + __popc(mask & lt_mask)] = (int16_t)m;
        nk += __popc(mask);
    }
    __syncwarp();
    for (int q = lane; q < nk; q += 32) {
        double ps = s_ps[s_order2[q]];
        if (ps != ps) ps = -1.0e300;
        int rank = 0;
        for (int q2 = 0; q2 < nk; ++q2) {
            double p2 = s_ps[s_order2[q2]];
            if (p2 != p2) p2 = -1.0e300;
            rank += (p2 > ps || (p2 == ps && q2 < q)) ? 1 : 0;
        }
        s_rank[q] = (int16_t)rank;
    }
    int off = 0;
    if (lane == 0) {
        off = atomicAdd(out_total, nk);
        out_offset[img] = off;
        out_count[img] = nk;
    }
    off = __shfl_sync(kFull, off, 0);
    __syncwarp();
    for (int e = lane; e < nk * C; e += 32) {
        const int q = e / C, c = e - q * C;
        const int dst = off + s_rank[q];
        if (dst >= capacity_rows) continue;
        const int row = s_order[s_order2[q]];
        const float4 v = xyvs[row * C + c];
        float *o = out_poses + ((size_t)dst * C + c) * OG_POSE_COLS;
        o[0] = unset_to_zero(v.x);
        o[1] = unset_to_zero(v.y);
        o[2] = unset_to_zero(v.z);
        o[3] = unset_to_zero(v.w);
        o[4] = unset_to_zero(score[row * C + c]);
        o[5] = unset_to_zero((float)ids[row * C + c]);
    }
}

GroupArgs to_args(const GroupLaunch &g) {
    GroupArgs a;
    a.C = g.c;
    a.L = g.l;
    a.K = g.k;
    a.sk = g.sk;
    a.dist_max = g.dist_max;
    a.use_scale = g.use_scale;
    a.person_thre = g.person_thre;
    a.sort_dim = g.sort_dim;
    a.smem_rows = g.smem_rows;
    a.slab = g.slab;
    a.slab_stride = g.slab_stride;
    return a;
}

}  // namespace

size_t group_smem_bytes(const GroupLaunch &g) { return make_layout(g.c, g.l, g.k, g.smem_rows).total; }

int group_warp_rows(const GroupLaunch &g) { return min(kWarpRowsMax, g.l * g.k); }
size_t group_warp_smem_bytes(const GroupLaunch &g) {
    return make_warp_layout(g.c, g.k, group_warp_rows(g)).total;
}

int prepare_group_kernel(size_t smem_bytes, size_t warp_smem_bytes) {
    OG_CUDA_TRY(cudaFuncSetAttribute(group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
    OG_CUDA_TRY(cudaFuncSetAttribute(group_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)warp_smem_bytes));
    return OG_OK;
}

int launch_group(const GroupLaunch &g, const float *limbs, float *out_poses, int capacity_rows,
                 int32_t *out_offset, int32_t *out_count, int32_t *out_total, cudaStream_t s) {
    if (g.n == 0) return OG_OK;
    const size_t smem = group_smem_bytes(g);
    if (g.needs_cta != nullptr) {
        // one warp per image first; images that outgrow its table are redone by the CTA kernel
        group_warp_kernel<<<g.n, 32, group_warp_smem_bytes(g), s>>>(
            to_args(g), group_warp_rows(g), limbs, out_poses, capacity_rows, out_offset, out_count,
            out_total, g.needs_cta);
        OG_CUDA_TRY(cudaGetLastError());
    }
    group_kernel<<<g.n, kGroupThreads, smem, s>>>(to_args(g), limbs, out_poses, capacity_rows,
                                                  out_offset, out_count, out_total, g.needs_cta);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

}  // namespace og
