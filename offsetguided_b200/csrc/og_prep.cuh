// Limb-row preparation shared by K2 (og_limbs.cu, fused behind the scoring of the same CTA) and
// the stand-alone prepare kernel of K3 (og_group.cu, for limb tables a caller supplies):
// the half of every grouping step that depends on the K rows of ONE (image, limb type) only —
//   * distance / border gate                              (reference decoder/group.py:64-76)
//   * sort by limb score, canonical stable order          (group.py:232)
//   * best row per distinct to-joint id                   (group.py:233-239)
// Outputs per (image, limb type):
//   prep[K + 1]   kept row indices in order, their count at [K]        (global-slab grouping kernel)
//   rec[3 * K]    the kept rows, compacted, 48 bytes each              (warp grouping kernel):
//                 {id1, id2 (int bits), limb score, -}, {x1, y1, v1, scale1}, {x2, y2, v2, scale2}
//   cnt           the count again, dense per (image, limb); bit 16 (kPrepDupFrom) is set when two
//                 kept rows start at the same from-joint id (never on the decode path, where K2
//                 makes one row per from-candidate; caller-supplied tables may)
#pragma once

#include "og_common.cuh"

namespace og {

constexpr int kPrepDupFrom = 1 << 16;
constexpr int kPrepCountMask = 0xffff;

// KMAX = the CTA size (32 / 64 / 128 >= K): the scratch is sized by it, because a CTA that asks
// for more than a few KB of shared memory cannot join an SM whose carve-out the streaming kernels
// of the next call (no shared memory at all) have set to "all L1" until that SM has drained.
template <int KMAX>
struct PrepShared {
    float4 rec[KMAX * 3];
    float sc[KMAX];
    int id2[KMAX];
    int sorted[KMAX];
    uint8_t valid[KMAX];
    uint8_t keep[KMAX];
};

// Every thread of the CTA calls this (blockDim.x >= K); thread tid < K owns limb row `r`
// (the 13 columns of decoder/collect.py:220-222).
template <int KMAX>
__device__ __forceinline__ void prepare_limb_rows(const float (&r)[OG_LIMB_COLS], int tid, int K,
                                                  float dist_max, int use_scale, PrepShared<KMAX> &sh,
                                                  int32_t *__restrict__ prep_out,
                                                  float4 *__restrict__ rec_out,
                                                  int32_t *__restrict__ cnt_out) {
    bool valid = false;
    float sc = 0.0f;
    if (tid < K) {
        // gate (group.py:64-76): distance test and both endpoints strictly inside the image
        float lim = dist_max;
        if (use_scale) lim = (r[12] != r[12]) ? r[12] : fmaxf(dist_max, r[12]);
        valid = (r[8] < lim) && (r[0] > 0.f) && (r[4] > 0.f) && (r[3] > 0.f) && (r[1] > 0.f) &&
                (r[10] == r[10]);
        sc = r[10];
        sh.sc[tid] = sc;
        sh.valid[tid] = valid ? 1 : 0;
    }
    const int nvalid = __syncthreads_count(valid);
    // rank = number of valid rows that precede this one: limb score desc, row asc
    // (group.py:232 with the canonical stable order)
    if (valid) {
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const float sj = sh.sc[j];
            rank += (sh.valid[j] && (sj > sc || (sj == sc && j < tid))) ? 1 : 0;
        }
        sh.sorted[rank] = tid;
        sh.id2[rank] = (int)r[7];
        sh.rec[rank * 3 + 0] = make_float4(__int_as_float((int)r[6]), __int_as_float((int)r[7]), sc, 0.0f);
        sh.rec[rank * 3 + 1] = make_float4(r[0], r[1], r[2], r[11]);
        sh.rec[rank * 3 + 2] = make_float4(r[3], r[4], r[5], r[12]);
    }
    __syncthreads();
    // keep the best row per distinct to-joint id (group.py:233-239)
    bool keep = false;
    if (tid < nvalid) {
        const int t = sh.id2[tid];
        keep = true;
        for (int r2 = 0; r2 < tid; ++r2) keep = keep && (sh.id2[r2] != t);
        sh.keep[tid] = keep ? 1 : 0;
    }
    const int kk = __syncthreads_count(keep);
    bool dup = false;
    if (keep) {
        int pos = 0;
        const float id1 = sh.rec[tid * 3].x;            // the int bits of the from-joint id
        for (int r2 = 0; r2 < tid; ++r2) {
            pos += sh.keep[r2];
            dup = dup || (sh.keep[r2] && __float_as_int(sh.rec[r2 * 3].x) == __float_as_int(id1));
        }
        prep_out[pos] = sh.sorted[tid];
        rec_out[pos * 3 + 0] = sh.rec[tid * 3 + 0];
        rec_out[pos * 3 + 1] = sh.rec[tid * 3 + 1];
        rec_out[pos * 3 + 2] = sh.rec[tid * 3 + 2];
    }
    const int any_dup = __syncthreads_or(dup);
    if (tid == 0) {
        prep_out[K] = kk;
        *cnt_out = kk | (any_dup ? kPrepDupFrom : 0);
    }
}

}  // namespace og
