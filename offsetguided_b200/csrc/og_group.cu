// K3 — greedy limb-ordered keypoint grouping (reference decoder/group.py:39-240, which
// runs per image in a CPU process pool).  Two kernels:
//
//   group_prepare_kernel   one CTA per (limb type, image): the part of every limb step that
//       depends on the limb rows only — distance / border gate (group.py:64-76), sort by limb
//       score (canonical stable order) and "best row per to-joint id" (group.py:222-240).
//       It writes the kept row indices in order; all L x N instances run in parallel, off the
//       sequential path.
//   group_kernel           one CTA per image walks the skeleton in the reference's limb order.
//
// The person table lives in shared memory as three planes per (person row, joint):
//   ids   int32   global keypoint id (-1 = unset)          pose column 5
//   score float   best limb score seen for the joint       pose column 4
//   xyvs  float4  x, y, keypoint score, keypoint scale     pose columns 0..3
// Persons are addressed through an `order` list (logical position -> table row), so
// deleting merged persons is a compaction of that small list, never a row move.
//
// The numpy program is sequential only in appearance; per limb type it is
//   (1) match every person against every kept limb   -> one thread per (person, limb) pair
//   (2) apply "both ends known" / "one end known"    -> one thread per person; the
//       reference's fancy-index scatter semantics (row-major pair order, last pair wins,
//       right-hand sides read before the statement) reduce to "the LAST kept limb that
//       matches person m" (atomicMax over the pair threads), applied by the thread owning m
//   (3) merge persons sharing exactly two ids        -> one thread per person pair; partner
//       = largest b (atomicMax); survivors read rows that are being deleted, which nobody
//       writes
//   (4) append unclaimed limbs as new persons, using the reference's column-sum rule
//       including its (-1)+(+1) cancellation quirk (group.py:166).
// The table has `smem_rows` rows; if an image needs more (never seen outside noise inputs)
// the same CTA restarts that image with the table in a global slab of L*K rows, which cannot
// overflow (every new person consumes one limb row).
#include "og_common.cuh"

#include <mutex>

namespace og {

namespace {

#ifndef OG_K3_THREADS
#define OG_K3_THREADS 256
#endif
constexpr int kGroupThreads = OG_K3_THREADS;
constexpr int kPrepThreads = 128;          // one thread per limb row, >= OG_MAX_TOPK
static_assert(kPrepThreads >= OG_MAX_TOPK, "prepare kernel needs one thread per limb row");

// Optional phase profile (build with -DOG_K3_PROFILE): thread 0 accumulates clock64()
// deltas per phase into og_k3_prof[phase]; og_k3_prof[15] counts CTAs.
#ifdef OG_K3_PROFILE
__device__ unsigned long long og_k3_prof[16];
#define OG_K3_PROF(ph)                                   \
    do {                                                 \
        if (threadIdx.x == 0) {                          \
            const long long t__ = clock64();             \
            prof_acc[ph] += t__ - prof_last;             \
            prof_last = t__;                             \
        }                                                \
    } while (0)
#else
#define OG_K3_PROF(ph) do { } while (0)
#endif

enum Flag { kAnyP1 = 0, kAnyP2, kAnyMerge, kOutOffset, kNNew, kNumFlags = 8 };

struct GroupArgs {
    int C, L, K;
    SkeletonDev sk;
    double person_thre;
    int sort_dim;
    int smem_rows;
    float *slab;     // per image: xyvs float4[PMAX*C], score float[PMAX*C], ids int[PMAX*C]
    size_t slab_stride;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// prepare: gate + sort + dedup of one limb type's K rows.
// prep[(image * L + limb) * (K + 1)] = kept row indices in order, then their count at [K].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kPrepThreads)
group_prepare_kernel(const float *__restrict__ limbs, int L, int K, float dist_max, int use_scale,
                     int32_t *__restrict__ prep) {
    __shared__ float s_conn[OG_MAX_TOPK * OG_LIMB_COLS];
    __shared__ float s_sc[OG_MAX_TOPK];
    __shared__ int s_id2[OG_MAX_TOPK];
    __shared__ int s_sorted[OG_MAX_TOPK];
    __shared__ uint8_t s_valid[OG_MAX_TOPK];
    __shared__ uint8_t s_keep[OG_MAX_TOPK];
    const int tid = threadIdx.x;
    const size_t inst = (size_t)blockIdx.y * L + blockIdx.x;          // (image, limb type)
    const float *src = limbs + inst * K * OG_LIMB_COLS;
    int32_t *out = prep + inst * (K + 1);
    for (int i = tid; i < K * OG_LIMB_COLS; i += kPrepThreads) s_conn[i] = __ldg(src + i);
    __syncthreads();
    // gate (group.py:64-76): distance test and both endpoints strictly inside the image
    bool valid = false;
    float sc = 0.0f;
    if (tid < K) {
        const float *r = s_conn + tid * OG_LIMB_COLS;
        float lim = dist_max;
        if (use_scale) lim = (r[12] != r[12]) ? r[12] : fmaxf(dist_max, r[12]);
        valid = (r[8] < lim) && (r[0] > 0.f) && (r[4] > 0.f) && (r[3] > 0.f) && (r[1] > 0.f) &&
                (r[10] == r[10]);
        sc = r[10];
        s_sc[tid] = sc;
        s_valid[tid] = valid ? 1 : 0;
    }
    const int nvalid = __syncthreads_count(valid);
    // rank = number of valid rows that precede this one: limb score desc, row asc
    // (group.py:232 with the canonical stable order)
    if (valid) {
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const float sj = s_sc[j];
            rank += (s_valid[j] && (sj > sc || (sj == sc && j < tid))) ? 1 : 0;
        }
        s_sorted[rank] = tid;
        s_id2[rank] = (int)s_conn[tid * OG_LIMB_COLS + 7];
    }
    __syncthreads();
    // keep the best row per distinct to-joint id (group.py:233-239)
    bool keep = false;
    if (tid < nvalid) {
        const int t = s_id2[tid];
        keep = true;
        for (int r2 = 0; r2 < tid; ++r2) keep = keep && (s_id2[r2] != t);
        s_keep[tid] = keep ? 1 : 0;
    }
    const int kk = __syncthreads_count(keep);
    if (keep) {
        int pos = 0;
        for (int r2 = 0; r2 < tid; ++r2) pos += s_keep[r2];
        out[pos] = s_sorted[tid];
    }
    if (tid == 0) out[K] = kk;
}

// ---------------------------------------------------------------------------
// grouping
// ---------------------------------------------------------------------------
struct Layout {
    size_t xyvs, score, ids, conn, rows, k_int, p_i32, p_i16, p_u8, p_f64, warp, flags, total;
};

// k_int: 6 int arrays of K; p_i32: 3 int arrays of PMAX; p_i16: 3 int16 arrays of PMAX
__host__ __device__ inline Layout make_layout(int C, int L, int K, int smem_rows) {
    const size_t pmax = (size_t)L * K;
    Layout lo;
    size_t at = 0;
    lo.xyvs = at;  at += (size_t)smem_rows * C * 16;
    lo.score = at; at += (size_t)smem_rows * C * 4;
    lo.ids = at;   at += (size_t)smem_rows * C * 4;
    lo.conn = at;  at += 2 * align_up((size_t)K * OG_LIMB_COLS * 4, 16);     // double buffer
    lo.rows = at;  at += 2 * align_up((size_t)(K + 1) * 4, 16);               // kept rows + count
    lo.k_int = at; at += (size_t)6 * K * 4;
    at = align_up(at, 8);
    lo.p_f64 = at; at += pmax * 8;
    lo.p_i32 = at; at += (size_t)3 * pmax * 4;
    lo.p_i16 = at; at += align_up((size_t)3 * pmax * 2, 4);
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

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gmem_src));
}

__global__ void __launch_bounds__(kGroupThreads)
group_kernel(GroupArgs a, const float *__restrict__ limbs, const int32_t *__restrict__ prep,
             float *__restrict__ out_poses, int capacity_rows, int32_t *__restrict__ out_offset,
             int32_t *__restrict__ out_count, int32_t *__restrict__ out_total) {
    extern __shared__ __align__(16) unsigned char smem[];
#ifdef OG_K3_PROFILE
    long long prof_acc[10] = {0};
    long long prof_last = clock64();
#endif
    const int C = a.C, L = a.L, K = a.K;
    const int pmax = L * K;
    const Layout lo = make_layout(C, L, K, a.smem_rows);
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int img = blockIdx.x;

    float *s_conn_buf[2];
    s_conn_buf[0] = reinterpret_cast<float *>(smem + lo.conn);
    s_conn_buf[1] = s_conn_buf[0] + align_up((size_t)K * OG_LIMB_COLS * 4, 16) / 4;
    int *s_rows_buf[2];                     // [0..K) kept rows in order, [K] their count
    s_rows_buf[0] = reinterpret_cast<int *>(smem + lo.rows);
    s_rows_buf[1] = s_rows_buf[0] + align_up((size_t)(K + 1) * 4, 16) / 4;
    int *kbase = reinterpret_cast<int *>(smem + lo.k_int);
    int *s_kind1 = kbase + 0 * K;
    int *s_kind2 = kbase + 1 * K;
    float *s_kscore = reinterpret_cast<float *>(kbase + 2 * K);
    int *s_n1 = kbase + 3 * K;
    int *s_n2 = kbase + 4 * K;
    int *s_new = kbase + 5 * K;
    int *p32 = reinterpret_cast<int *>(smem + lo.p_i32);
    int *s_pk1 = p32 + 0 * pmax;
    int *s_pk2 = p32 + 1 * pmax;
    int *s_blast = p32 + 2 * pmax;
    int16_t *pbase = reinterpret_cast<int16_t *>(smem + lo.p_i16);
    int16_t *s_order = pbase + 0 * pmax;
    int16_t *s_order2 = pbase + 1 * pmax;
    int16_t *s_pos = pbase + 2 * pmax;
    uint8_t *s_del = smem + lo.p_u8;
    double *s_ps = reinterpret_cast<double *>(smem + lo.p_f64);
    int *s_warp = reinterpret_cast<int *>(smem + lo.warp);
    volatile int *s_flag = reinterpret_cast<volatile int *>(smem + lo.flags);

    const float *limbs_img = limbs + (size_t)img * L * K * OG_LIMB_COLS;
    const int32_t *prep_img = prep + (size_t)img * L * (K + 1);

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

        // rows and kept list of limb type li + 1 are fetched with cp.async while type li runs
        auto fetch_rows = [&](int li_next) {
            float *dst = s_conn_buf[li_next & 1];
            const float *src = limbs_img + (size_t)li_next * K * OG_LIMB_COLS;
            for (int i = tid; i < K * OG_LIMB_COLS; i += T) cp_async4(dst + i, src + i);
            int *rdst = s_rows_buf[li_next & 1];
            const int32_t *rsrc = prep_img + (size_t)li_next * (K + 1);
            for (int i = tid; i <= K; i += T) cp_async4(rdst + i, rsrc + i);
            asm volatile("cp.async.commit_group;");
        };
        fetch_rows(0);
        for (int li = 0; li < L && !overflow; ++li) {
            const int jf = a.sk.from[li], jt = a.sk.to[li];
            const float *s_conn = s_conn_buf[li & 1];
            const int *s_rows = s_rows_buf[li & 1];
            asm volatile("cp.async.wait_all;" ::: "memory");        // data of type li has landed
            if (tid < kNumFlags) s_flag[tid] = 0;
            __syncthreads();
            // The buffers of type li + 1 are those of type li - 1: only after the barrier has every
            // thread finished reading them (a type with kk == 0 leaves its iteration without one).
            if (li + 1 < L) fetch_rows(li + 1);
            OG_K3_PROF(0);
            const int kk = s_rows[K];
            if (kk == 0) continue;                                         // group.py:84-85

            // ---- stage ids / scores of the kept rows; reset the per-person match slots
            for (int j = tid; j < kk; j += T) {
                const float *r = s_conn + s_rows[j] * OG_LIMB_COLS;
                s_kind1[j] = (int)r[6];
                s_kind2[j] = (int)r[7];
                s_kscore[j] = r[10];
                s_n1[j] = 0;
                s_n2[j] = 0;
            }
            for (int m = tid; m < mm; m += T) {
                s_pk1[m] = -1;
                s_pk2[m] = -1;
                s_blast[m] = -1;
                s_del[m] = 0;
            }
            __syncthreads();
            OG_K3_PROF(1);
            // ---- match persons x kept limbs on the pre-update snapshot (group.py:87-109),
            //      one thread per pair; the last matching limb of a person wins (atomicMax)
            for (int e = tid; e < mm * kk; e += T) {
                const int m = e / kk, j = e - m * kk;
                const int row = s_order[m];
                const int ms = (ids[row * C + jf] == s_kind1[j] ? 1 : 0) +
                               (ids[row * C + jt] == s_kind2[j] ? 1 : 0);
                if (ms == 0) continue;
                const float sc = s_kscore[j];
                const bool rep = (sc > score[row * C + jt]) || (sc > score[row * C + jf]);
                if (ms == 2) {
                    atomicAdd(&s_n2[j], 1);
                    if (rep) {
                        atomicMax(&s_pk2[m], j);
                        s_flag[kAnyP2] = 1;
                    }
                } else {
                    atomicAdd(&s_n1[j], 1);
                    if (rep) {
                        atomicMax(&s_pk1[m], j);
                        s_flag[kAnyP1] = 1;
                    }
                }
            }
            __syncthreads();
            OG_K3_PROF(2);
            // ---- apply: both ends known (group.py:114-119), then one end known (:124-135)
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
                    const float *r = s_conn + s_rows[k1] * OG_LIMB_COLS;
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
            OG_K3_PROF(3);
            // ---- merge persons sharing exactly two keypoint ids (group.py:140-155), one thread
            //      per ordered pair p < q; the largest partner of p wins
            int mm_after = mm;
            if (mm >= 2) {
                for (int e = tid; e < mm * mm; e += T) {
                    const int p = e / mm, q = e - p * mm;
                    if (q <= p) continue;
                    const int rowa = s_order[p], rowb = s_order[q];
                    int cnt = 0;
                    for (int c = 0; c < C; ++c) {
                        const int ia = ids[rowa * C + c];
                        cnt += (ia != -1 && ia == ids[rowb * C + c]) ? 1 : 0;
                    }
                    if (cnt == 2) {
                        atomicMax(&s_blast[p], q);
                        s_del[q] = 1;
                        s_flag[kAnyMerge] = 1;
                    }
                }
                __syncthreads();
                OG_K3_PROF(4);
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
            // ---- warp 0: unclaimed limbs start new persons (group.py:166-177), column-sum
            //      rule with its (-1) + (+1) cancellation
            if (wid == 0) {
                const int w2 = s_flag[kAnyP2] ? -1 : 2;
                const int w1 = s_flag[kAnyP1] ? -1 : 1;
                int nnew = 0;
                for (int j0 = 0; j0 < kk; j0 += 32) {
                    const int j = j0 + lane;
                    const bool isnew = j < kk && (s_n2[j] * w2 + s_n1[j] * w1 == 0);
                    const unsigned mask = __ballot_sync(0xffffffffu, isnew);
                    if (isnew) s_new[nnew + __popc(mask & ((1u << lane) - 1u))] = j;
                    nnew += __popc(mask);
                }
                if (lane == 0) s_flag[kNNew] = nnew;
            }
            __syncthreads();
            OG_K3_PROF(5);
            const int nnew = s_flag[kNNew];
            if (nalloc + nnew > pcap) {
                overflow = true;          // uniform: restart this image on the global slab
                continue;
            }
            for (int e = tid; e < nnew * C; e += T) {
                const int q = e / C, c = e - q * C;
                const int j = s_new[q];
                const int at = (nalloc + q) * C + c;
                const float *r = s_conn + s_rows[j] * OG_LIMB_COLS;
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
            OG_K3_PROF(6);
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
        s_blast[q] = rank;
    }
    if (tid == 0) {
        const int off = atomicAdd(out_total, nk);
        s_flag[kOutOffset] = off;
        out_offset[img] = off;
        out_count[img] = nk;
    }
    __syncthreads();
    OG_K3_PROF(7);
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
#ifdef OG_K3_PROFILE
    OG_K3_PROF(8);
    if (threadIdx.x == 0) {
        for (int ph = 0; ph < 10; ++ph) atomicAdd(&og_k3_prof[ph], (unsigned long long)prof_acc[ph]);
        atomicAdd(&og_k3_prof[15], 1ull);
    }
#endif
}

GroupArgs to_args(const GroupLaunch &g) {
    GroupArgs a;
    a.C = g.c;
    a.L = g.l;
    a.K = g.k;
    a.sk = g.sk;
    a.person_thre = g.person_thre;
    a.sort_dim = g.sort_dim;
    a.smem_rows = g.smem_rows;
    a.slab = g.slab;
    a.slab_stride = g.slab_stride;
    return a;
}

}  // namespace

int read_k3_profile(unsigned long long *out16, bool reset) {
#ifdef OG_K3_PROFILE
    OG_CUDA_TRY(cudaMemcpyFromSymbol(out16, og_k3_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long zeros[16] = {0};
        OG_CUDA_TRY(cudaMemcpyToSymbol(og_k3_prof, zeros, sizeof(zeros)));
    }
    return OG_OK;
#else
    (void)out16;
    (void)reset;
    set_error("library built without -DOG_K3_PROFILE");
    return OG_ERR_UNSUPPORTED;
#endif
}

size_t group_smem_bytes(const GroupLaunch &g) { return make_layout(g.c, g.l, g.k, g.smem_rows).total; }

size_t group_prep_ints(const GroupLaunch &g) { return (size_t)g.n * g.l * (g.k + 1); }

// The attribute belongs to the kernel (per device), not to a handle: handles with different
// table sizes coexist, so it is only ever raised.
int prepare_group_kernel(size_t smem_bytes) {
    static std::mutex mu;
    static size_t granted[64] = {0};
    int dev = 0;
    OG_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && smem_bytes <= granted[dev]) return OG_OK;
    OG_CUDA_TRY(cudaFuncSetAttribute(group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
    if (dev >= 0 && dev < 64) granted[dev] = smem_bytes;
    return OG_OK;
}

int launch_group(const GroupLaunch &g, const float *limbs, float *out_poses, int capacity_rows,
                 int32_t *out_offset, int32_t *out_count, int32_t *out_total, cudaStream_t s) {
    if (g.n == 0) return OG_OK;
    group_prepare_kernel<<<dim3(g.l, g.n), kPrepThreads, 0, s>>>(limbs, g.l, g.k, g.dist_max,
                                                                 g.use_scale, g.prep);
    OG_CUDA_TRY(cudaGetLastError());
    group_kernel<<<g.n, kGroupThreads, group_smem_bytes(g), s>>>(
        to_args(g), limbs, g.prep, out_poses, capacity_rows, out_offset, out_count, out_total);
    OG_CUDA_TRY(cudaGetLastError());
    return OG_OK;
}

}  // namespace og
