// a10 — uniform negative sampling on the device (data/dataloader.py:56-57, 92-129 of the reference).
//
// Semantics = the reference's collate loop: every flagged slot of a row draws a uniform item in [0, n_items); then the
// row is re-checked as a whole — a slot is flagged if its item is one of the user's TRAIN items (CSR row, binary
// search) or, with distinct_in_row, if a HIGHER slot of the same row holds the same item (numpy's
// `np.isin(..., assume_unique=True)` sort path flags every duplicate but the last occurrence) — until no
// slot is flagged.  Randomness: Philox4x32-10, key = (seed_lo, seed_hi ^ step_hi), counter = (b, j, q / 4, step_lo) where
// q counts the 32-bit words slot (b, j) has consumed; word q % 4 of the block is used.  A word r becomes an item by
// Lemire's multiply-shift with rejection (exactly uniform).  The index stream is therefore a pure function of
// (seed, step, b, j) — oracle/philox.py restates it in numpy and the parity test is bit-exact.
// One warp per batch row; the row's candidate items live in shared memory.
#include "hsk_topk.cuh"

namespace hsk {

constexpr int kSamplerMaxN = 1024;
constexpr int kSamplerWarps = 4;
constexpr int kSamplerMaxRounds = 256;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 'popular' strategy (data/dataloader.py:59-64): item ~ pop^alpha.  The caller passes the distribution as a 64-bit
// fixed-point CDF (cdf[i] = floor(2^64 * sum_{t <= i} p_t), last entry 2^64 - 1); a draw takes two Philox words r64 and
// returns the first i with cdf[i] > r64 — integer arithmetic only, so the numpy oracle reproduces it bit for bit.
__device__ __forceinline__ uint32_t draw_item_cdf(uint32_t b, uint32_t j, uint32_t& q, uint32_t step_lo, uint32_t k0, uint32_t k1,
                                                  const uint64_t* __restrict__ cdf, uint32_t n_items) {
    uint32_t w[4];
    uint32_t hi, lo;
    philox4x32_10(b, j, q >> 2, step_lo, k0, k1, w);
    hi = w[q & 3u];
    ++q;
    if ((q & 3u) == 0u) philox4x32_10(b, j, q >> 2, step_lo, k0, k1, w);
    lo = w[q & 3u];
    ++q;
    const uint64_t r = ((uint64_t)hi << 32) | lo;
    uint32_t a = 0, n = n_items;   // upper_bound
    while (n > 0) {
        const uint32_t half = n >> 1;
        if (__ldg(cdf + a + half) <= r) { a += half + 1; n -= half + 1; } else { n = half; }
    }
    return a < n_items ? a : n_items - 1;
}

// next uniform item of slot (b, j); q = words consumed so far (updated)
__device__ __forceinline__ uint32_t draw_item(uint32_t b, uint32_t j, uint32_t& q, uint32_t step_lo, uint32_t k0, uint32_t k1,
                                              uint32_t n_items) {
    const uint32_t thresh = (0u - n_items) % n_items;  // 2^32 mod n_items
    for (;;) {
        uint32_t w[4];
        philox4x32_10(b, j, q >> 2, step_lo, k0, k1, w);
        const uint32_t r = w[q & 3u];
        ++q;
        const uint64_t m = (uint64_t)r * (uint64_t)n_items;
        if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
    }
}

__global__ void __launch_bounds__(kSamplerWarps * 32) sample_negatives_kernel(
    const int64_t* __restrict__ u_idx, const int64_t* __restrict__ pos_idx, int B, int N, uint32_t n_items, int64_t n_users,
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, uint32_t k0, uint32_t k1, uint32_t step_lo,
    int distinct, const uint64_t* __restrict__ cdf, int64_t* __restrict__ i_idx, int32_t* status) {
    extern __shared__ uint32_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kSamplerWarps + warp;
    if (b >= B) return;
    uint32_t* val = sm + (size_t)warp * 2 * N;
    uint32_t* cnt = val + N;
    const int64_t u = u_idx[b];
    int64_t* out = i_idx + (int64_t)b * (N + 1);
    if (bad_index(u, n_users)) {
        if (lane == 0 && status) atomicOr(status, HSK_STATUS_BAD_INDEX);
        for (int j = lane; j <= N; j += 32) out[j] = 0;
        return;
    }
    const int64_t lo = indptr[u], hi = indptr[u + 1];
    for (int j = lane; j < N; j += 32) { cnt[j] = 0u; val[j] = 0xFFFFFFFFu; }  // 0xFFFFFFFF = "flagged, must draw"
    __syncwarp();
    bool clean = false;
    for (int round = 0; round < kSamplerMaxRounds; ++round) {
        // draw for flagged slots (flag is kept in the top bit of cnt)
        for (int j = lane; j < N; j += 32) {
            if (round == 0 || (cnt[j] & 0x80000000u)) {
                uint32_t q = cnt[j] & 0x7FFFFFFFu;
                const uint32_t it = cdf ? draw_item_cdf((uint32_t)b, (uint32_t)j, q, step_lo, k0, k1, cdf, n_items)
                                        : draw_item((uint32_t)b, (uint32_t)j, q, step_lo, k0, k1, n_items);
                val[j] = it;
                cnt[j] = q | 0x40000000u;  // bit 30 = "drawn this round"
            }
        }
        __syncwarp();
        bool any = false;
        for (int j = lane; j < N; j += 32) {
            const uint32_t c = cnt[j];
            const uint32_t it = val[j];
            bool f = false;
            if (c & 0x40000000u) f = csr_contains(indices, lo, hi, (int32_t)it);
            if (!f && distinct) {
                for (int j2 = j + 1; j2 < N; ++j2)
                    if (val[j2] == it) { f = true; break; }
            }
            cnt[j] = (c & 0x3FFFFFFFu) | (f ? 0x80000000u : 0u);
            any |= f;
        }
        __syncwarp();
        if (!__any_sync(kFull, any)) { clean = true; break; }
    }
    // round cap hit (e.g. more distinct negatives asked for than the user has non-train items): the row still holds flagged
    // slots — report it instead of emitting them silently
    if (!clean && lane == 0 && status) atomicOr(status, HSK_STATUS_SAMPLER_ROUNDS);
    if (lane == 0) out[0] = pos_idx ? pos_idx[b] : 0;
    for (int j = lane; j < N; j += 32) out[1 + j] = (int64_t)val[j];
}

}  // namespace hsk

using namespace hsk;

extern "C" int hsk_sample_negatives(const int64_t* u_idx, const int64_t* pos_idx, int B, int N, int64_t n_items,
                                    int64_t n_users, const int64_t* csr_indptr, const int32_t* csr_indices, uint64_t seed,
                                    uint64_t step, int distinct_in_row, const uint64_t* pop_cdf, int64_t* i_idx, int32_t* status,
                                    hsk_stream_t stream) {
    HSK_REQUIRE(u_idx && csr_indptr && csr_indices && i_idx, "hsk_sample_negatives: null pointer");
    HSK_REQUIRE(B >= 0 && N >= 1 && N <= kSamplerMaxN, "hsk_sample_negatives: need 1 <= N <= %d (N=%d)", kSamplerMaxN, N);
    HSK_REQUIRE(n_items >= 1 && n_items < 0x7FFFFFFFll, "hsk_sample_negatives: n_items must fit int32");
    if (B == 0) return HSK_OK;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32);
    const size_t smem = (size_t)kSamplerWarps * 2 * N * sizeof(uint32_t);
    sample_negatives_kernel<<<(B + kSamplerWarps - 1) / kSamplerWarps, kSamplerWarps * 32, smem, as_stream(stream)>>>(
        u_idx, pos_idx, B, N, (uint32_t)n_items, n_users, csr_indptr, csr_indices, k0, k1, (uint32_t)step, distinct_in_row,
        pop_cdf, i_idx, status);
    return check_launch("hsk_sample_negatives");
}
