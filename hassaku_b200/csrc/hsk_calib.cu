// f3 — the recommendation distributions of FullEvaluatorCalibrationDecorator (eval/eval.py:174-179 in the reference):
//   q_k[b, :] = mean over the first k ranked items of item_tag_mtx[top_ids[b, j], :]      for every k in CALIBRATION_K_VALUES
// The reference gathers item_tag_mtx[top ids] into a [B, k_max, T] tensor per batch and sums slices of it; here one warp
// walks a user's ranked list once, lanes own tags, the running sums are emitted at every requested k: the [B, k_max, T]
// intermediate (151 MB per 18 944-user batch at T = 20) never exists.  Sums are carried in fp64 and rounded once.
#include "hsk_common.cuh"

namespace hsk {

constexpr int kMaxCalibKs = 8;
struct CalibKs {
    int n;
    int k[kMaxCalibKs];
};

__global__ void __launch_bounds__(128) topk_tag_means_kernel(const int32_t* __restrict__ top_ids, int B, int k_list,
                                                             const float* __restrict__ item_tag, int64_t n_items, int T, CalibKs ks,
                                                             float* __restrict__ out, int32_t* status) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int32_t* ids = top_ids + (int64_t)warp * k_list;
    int k_max = 0;
    for (int t = 0; t < ks.n; ++t) k_max = max(k_max, ks.k[t]);
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int tag = t0 + lane;
        double acc = 0.0;
        for (int j = 0; j < k_max; ++j) {
            const int32_t it = ids[j];            // warp-uniform
            if (it >= 0) {
                if ((int64_t)it >= n_items) {
                    if (lane == 0 && status) atomicOr(status, HSK_STATUS_BAD_INDEX);
                } else if (tag < T) {
                    acc += (double)__ldg(item_tag + (int64_t)it * T + tag);
                }
            }
            for (int t = 0; t < ks.n; ++t)
                if (j + 1 == ks.k[t] && tag < T) out[((int64_t)warp * ks.n + t) * T + tag] = (float)(acc / (double)ks.k[t]);
        }
    }
}

}  // namespace hsk

extern "C" int hsk_topk_tag_means(const int32_t* top_ids, int B, int k_list, const float* item_tag, int64_t n_items, int T,
                                  const int* ks, int n_ks, float* out, int32_t* status, hsk_stream_t stream) {
    HSK_REQUIRE(top_ids && item_tag && ks && out, "hsk_topk_tag_means: null pointer");
    HSK_REQUIRE(B >= 0 && k_list >= 1 && T >= 1 && n_items >= 1, "hsk_topk_tag_means: bad sizes");
    HSK_REQUIRE(n_ks >= 1 && n_ks <= hsk::kMaxCalibKs, "hsk_topk_tag_means: 1..%d values of k", hsk::kMaxCalibKs);
    hsk::CalibKs c;
    c.n = n_ks;
    for (int t = 0; t < n_ks; ++t) {
        HSK_REQUIRE(ks[t] >= 1 && ks[t] <= k_list, "hsk_topk_tag_means: k = %d outside the ranked list (%d entries)", ks[t], k_list);
        c.k[t] = ks[t];
    }
    if (B == 0) return HSK_OK;
    const int wpb = 4;
    hsk::topk_tag_means_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, hsk::as_stream(stream)>>>(top_ids, B, k_list, item_tag, n_items, T, c,
                                                                                          out, status);
    return hsk::check_launch("hsk_topk_tag_means");
}
