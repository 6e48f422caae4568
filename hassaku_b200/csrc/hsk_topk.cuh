// Top-k machinery shared by the evaluator kernels.
//
// A candidate is ONE 64-bit key:  (orderable(score) << 32) | (0xFFFFFFFF - item_id)
// so that a plain unsigned compare orders by score descending and, among equal scores, by item id ascending —
// a strict total order, which makes the result independent of tile / split / shard processing order (the
// reference's torch.topk leaves the order of ties implementation-defined).  Key 0 = empty slot.
//
// Per row a list of up to kCap = 256 keys is kept; when it may overflow it is sorted by one warp with a
// register-resident bitonic network (8 keys per lane) and cut back to the best k.
#pragma once
#include "hsk_common.cuh"

namespace hsk {

constexpr int kCap = 256;          // candidate-list capacity per row
constexpr int kKeysPerLane = 8;    // kCap / 32
constexpr int kMaxK = 128;         // largest supported k (K_VALUES max is 100, eval/eval.py:20)

__device__ __forceinline__ uint32_t orderable(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t o) {
    uint32_t u = o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
    return (static_cast<uint64_t>(orderable(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - id);
}
__device__ __forceinline__ float key_score(uint64_t key) { return from_orderable(static_cast<uint32_t>(key >> 32)); }
__device__ __forceinline__ int32_t key_id(uint64_t key) {
    return key == 0 ? -1 : static_cast<int32_t>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFu));
}

// Sort 256 keys held by a warp (element e = r * 32 + lane lives in key[r] of `lane`) in DESCENDING order.
__device__ __forceinline__ void warp_sort_desc(uint64_t (&key)[kKeysPerLane], int lane) {
#pragma unroll
    for (int k = 2; k <= kCap; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < kKeysPerLane; ++r) {
                    const int rp = r ^ jr;
                    if (rp > r) {
                        const bool desc = (((r * 32 + lane) & k) == 0);
                        const uint64_t a = key[r], b = key[rp];
                        const bool sw = desc ? (a < b) : (a > b);
                        key[r] = sw ? b : a;
                        key[rp] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < kKeysPerLane; ++r) {
                    const uint64_t other = __shfl_xor_sync(kFull, key[r], j);
                    const bool lower = (lane & j) == 0;
                    const bool desc = (((r * 32 + lane) & k) == 0);
                    const bool keep_max = (lower == desc);
                    const uint64_t mx = key[r] > other ? key[r] : other;
                    const uint64_t mn = key[r] > other ? other : key[r];
                    key[r] = keep_max ? mx : mn;
                }
            }
        }
    }
}

// element #pos (0 <= pos < 256) of a warp-held sorted list, broadcast to every lane
__device__ __forceinline__ uint64_t warp_list_at(const uint64_t (&key)[kKeysPerLane], int pos) {
    uint64_t v = 0;
#pragma unroll
    for (int r = 0; r < kKeysPerLane; ++r)
        if (r == (pos >> 5)) v = key[r];
    return __shfl_sync(kFull, v, pos & 31);
}

// membership test in a sorted int32 range [lo, hi) (one CSR row): lower_bound + compare
__device__ __forceinline__ bool csr_contains(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int32_t x) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(idx + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(idx + lo) == x;
}

// Cut a row's candidate list (n <= kCap keys at `list`, global or shared memory) back to its best k, sorted.
// Returns the new length min(n, k); *thr_key receives the k-th key (0 while the list holds fewer than k).
__device__ __forceinline__ int warp_prune_list(uint64_t* list, int n, int k, int lane, uint64_t* thr_key) {
    uint64_t key[kKeysPerLane];
#pragma unroll
    for (int r = 0; r < kKeysPerLane; ++r) {
        const int e = r * 32 + lane;
        key[r] = (e < n) ? list[e] : 0ull;
    }
    warp_sort_desc(key, lane);
#pragma unroll
    for (int r = 0; r < kKeysPerLane; ++r) {
        const int e = r * 32 + lane;
        if (e < k) list[e] = key[r];
    }
    *thr_key = (n >= k) ? warp_list_at(key, k - 1) : 0ull;
    return n < k ? n : k;
}

// Same, applying the exclusion mask lazily: entries [n_checked, n) have not been tested against the user's exclusion row
// yet (the scoring epilogues append raw candidates so that their hot loop has no dependent global loads); an excluded
// item is re-keyed to score -inf (eval/eval.py:250-251) before the sort.  The 8 binary searches of a lane are
// independent, so their latencies overlap.
__device__ __forceinline__ int warp_prune_list_masked(uint64_t* list, int n, int n_checked, int k, int lane, uint64_t* thr_key,
                                                      const int32_t* __restrict__ excl, int64_t lo, int64_t hi) {
    uint64_t key[kKeysPerLane];
#pragma unroll
    for (int r = 0; r < kKeysPerLane; ++r) {
        const int e = r * 32 + lane;
        key[r] = (e < n) ? list[e] : 0ull;
    }
    if (hi > lo) {
#pragma unroll
        for (int r = 0; r < kKeysPerLane; ++r) {
            const int e = r * 32 + lane;
            if (e >= n_checked && e < n) {
                const int32_t id = key_id(key[r]);
                if (csr_contains(excl, lo, hi, id)) key[r] = make_key(-INFINITY, (uint32_t)id);
            }
        }
    }
    warp_sort_desc(key, lane);
#pragma unroll
    for (int r = 0; r < kKeysPerLane; ++r) {
        const int e = r * 32 + lane;
        if (e < k) list[e] = key[r];
    }
    *thr_key = (n >= k) ? warp_list_at(key, k - 1) : 0ull;
    return n < k ? n : k;
}

// hsk_eval.cu: merge of n_lists sorted key lists per row laid out [list][row][kCap] (split plans of the eval kernels)
int launch_merge_keys(const uint64_t* lists, int n_lists, int rows, int k, float* out_scores, int32_t* out_ids, cudaStream_t s,
                      const float* Ub, const float* Gb, const int64_t* u_idx, int64_t n_users);

}  // namespace hsk
