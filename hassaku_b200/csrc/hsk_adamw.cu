// a8 — torch.optim.AdamW(params, lr, weight_decay).step as ONE streaming pass (train/trainer.py:52-53,147).
//
// HBM-bound: 28 B/element (read p,m,v,g; write p,m,v) + 4 B/element when the gradient is zeroed in the same
// pass (replaces optimizer.zero_grad(), trainer.py:148).  128-bit loads/stores, grid = k * SM count.
//
// Arithmetic orders (every op is a separately rounded fp32 op; scalars are python doubles rounded to fp32
// exactly where torch rounds them):
//   arith 0 — torch CUDA `_multi_tensor_adam` (torch/optim/adam.py:554-800) = the foreach kernels
//       p  = p * f(1 - lr*wd)                              _foreach_mul_
//       m  = fma(f(1-b1), g - m, m)                        _foreach_lerp_  (weight < 0.5 branch, contracted)
//       v  = v * f(b2);  v = fma(f(1-b2), g*g, v)          _foreach_mul_, _foreach_addcmul_ (a + s*(b*c))
//       dn = sqrt(v) / f(sqrt(bc2)) + f(eps)               _foreach_sqrt, _foreach_div_, _foreach_add_
//       p  = fma(f(-lr/bc1), m / dn, p)                    _foreach_addcdiv_ (a + s*(b/c))
//   arith 1 — torch CPU `_single_tensor_adam` (torch/optim/adam.py:347-547), vectorised ATen CPU kernels
//       v  = fma(f(1-b2)*g, g, v*f(b2));   p = p + (f(-lr/bc1)*m)/dn      (rest as above)
// Both are verified bitwise on the GPU box (tests/test_gpu_adamw.py) against torch.optim.AdamW itself.
#include "hsk_common.cuh"

namespace hsk {

struct AdamConsts {
    float decay;      // f(1 - lr*wd)      (AdamW)   | unused (Adam)
    float wd;         // f(wd)             (Adam L2)
    float w1;         // f(1 - beta1)
    float beta2;      // f(beta2)
    float w2;         // f(1 - beta2)
    float sqrt_bc2;   // f(sqrt(1 - beta2^t))
    float eps;        // f(eps)
    float step_size;  // f(-(lr / (1 - beta1^t)))
};

template <int ARITH, bool L2, bool DECAY>
__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g, const AdamConsts& c) {
    if (L2) g = (ARITH == 0) ? __fmaf_rn(c.wd, p, g) : __fmaf_rn(p, c.wd, g);  // grad.add(param, alpha=wd)
    if (DECAY) p = __fmul_rn(p, c.decay);
    m = __fmaf_rn(c.w1, __fsub_rn(g, m), m);
    if (ARITH == 0) {
        v = __fmul_rn(v, c.beta2);
        v = __fmaf_rn(c.w2, __fmul_rn(g, g), v);
    } else {
        v = __fmaf_rn(__fmul_rn(c.w2, g), g, __fmul_rn(v, c.beta2));
    }
    float dn = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.sqrt_bc2), c.eps);
    if (ARITH == 0) {
        p = __fmaf_rn(c.step_size, __fdiv_rn(m, dn), p);
    } else {
        p = __fadd_rn(p, __fdiv_rn(__fmul_rn(c.step_size, m), dn));
    }
}

template <int ARITH, bool L2, bool DECAY, bool ZERO>
__global__ void __launch_bounds__(256) adamw_dense_kernel(float* __restrict__ p, float* __restrict__ m,
                                                          float* __restrict__ v, float* __restrict__ g, int64_t n,
                                                          AdamConsts c) {
    const int64_t n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* g4 = reinterpret_cast<float4*>(g);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = p4[i], M = m4[i], V = v4[i], G = __ldcs(g4 + i);
        adam_elem<ARITH, L2, DECAY>(P.x, M.x, V.x, G.x, c);
        adam_elem<ARITH, L2, DECAY>(P.y, M.y, V.y, G.y, c);
        adam_elem<ARITH, L2, DECAY>(P.z, M.z, V.z, G.z, c);
        adam_elem<ARITH, L2, DECAY>(P.w, M.w, V.w, G.w, c);
        p4[i] = P;
        m4[i] = M;
        v4[i] = V;
        if (ZERO) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // scalar tail (n % 4 elements)
    const int64_t tail0 = n4 << 2;
    const int64_t t = tail0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        float P = p[t], M = m[t], V = v[t], G = g[t];
        adam_elem<ARITH, L2, DECAY>(P, M, V, G, c);
        p[t] = P;
        m[t] = M;
        v[t] = V;
        if (ZERO) g[t] = 0.f;
    }
}

// Same update with the step-dependent scalars read from DEVICE memory, so that the launch can be captured once in a CUDA
// graph and replayed for every step (the host rewrites the 8 floats before each replay).
template <int ARITH, bool L2, bool DECAY, bool ZERO>
__global__ void __launch_bounds__(256) adamw_dense_devc_kernel(float* __restrict__ p, float* __restrict__ m,
                                                               float* __restrict__ v, float* __restrict__ g, int64_t n,
                                                               const AdamConsts* __restrict__ cp) {
    const AdamConsts c = *cp;
    const int64_t n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* g4 = reinterpret_cast<float4*>(g);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = p4[i], M = m4[i], V = v4[i], G = __ldcs(g4 + i);
        adam_elem<ARITH, L2, DECAY>(P.x, M.x, V.x, G.x, c);
        adam_elem<ARITH, L2, DECAY>(P.y, M.y, V.y, G.y, c);
        adam_elem<ARITH, L2, DECAY>(P.z, M.z, V.z, G.z, c);
        adam_elem<ARITH, L2, DECAY>(P.w, M.w, V.w, G.w, c);
        p4[i] = P; m4[i] = M; v4[i] = V;
        if (ZERO) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        float P = p[t], M = m[t], V = v[t], G = g[t];
        adam_elem<ARITH, L2, DECAY>(P, M, V, G, c);
        p[t] = P; m[t] = M; v[t] = V;
        if (ZERO) g[t] = 0.f;
    }
}

static void fill_consts(AdamConsts& c, double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step) {
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    c.decay = (float)(1.0 - lr * weight_decay);
    c.wd = (float)weight_decay;
    c.w1 = (float)(1.0 - beta1);
    c.beta2 = (float)beta2;
    c.w2 = (float)(1.0 - beta2);
    c.sqrt_bc2 = (float)sqrt(bc2);
    c.eps = (float)eps;
    c.step_size = (float)(-(lr / bc1));
}

}  // namespace hsk

// host helper: the 8 fp32 scalars of one step (what hsk_adamw_dense computes internally), for hsk_adamw_dense_graph
extern "C" int hsk_adamw_consts(double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                                float* out8 /* HOST */) {
    HSK_REQUIRE(out8 && step >= 1, "hsk_adamw_consts: bad arguments");
    hsk::AdamConsts c;
    hsk::fill_consts(c, lr, beta1, beta2, eps, weight_decay, step);
    memcpy(out8, &c, sizeof(c));
    return HSK_OK;
}

extern "C" int hsk_adamw_dense_graph(float* p, float* m, float* v, float* g, int64_t n, const float* consts_dev,
                                     int decoupled_decay, int adam_l2, int zero_grad, hsk_stream_t stream) {
    using namespace hsk;
    HSK_REQUIRE(p && m && v && g && consts_dev, "hsk_adamw_dense_graph: null pointer");
    HSK_REQUIRE(aligned16(p) && aligned16(m) && aligned16(v) && aligned16(g) && aligned16(consts_dev), "hsk_adamw_dense_graph: pointers must be 16-byte aligned");
    if (n <= 0) return HSK_OK;
    const int threads = 256;
    int64_t want = ((n >> 2) + threads - 1) / threads;
    if (want < 1) want = 1;
    const int64_t cap = (int64_t)sm_count() * 16;
    const int blocks = (int)(want < cap ? want : cap);
    cudaStream_t s = as_stream(stream);
    const AdamConsts* cp = reinterpret_cast<const AdamConsts*>(consts_dev);
#define HSK_LAUNCH_ADAMG(L, D, Z) adamw_dense_devc_kernel<0, L, D, Z><<<blocks, threads, 0, s>>>(p, m, v, g, n, cp)
    if (adam_l2) { if (zero_grad) HSK_LAUNCH_ADAMG(true, false, true); else HSK_LAUNCH_ADAMG(true, false, false); }
    else if (decoupled_decay) { if (zero_grad) HSK_LAUNCH_ADAMG(false, true, true); else HSK_LAUNCH_ADAMG(false, true, false); }
    else { if (zero_grad) HSK_LAUNCH_ADAMG(false, false, true); else HSK_LAUNCH_ADAMG(false, false, false); }
    return check_launch("hsk_adamw_dense_graph");
}

extern "C" int hsk_adamw_dense(float* p, float* m, float* v, float* g, int64_t n, double lr, double beta1,
                               double beta2, double eps, double weight_decay, int64_t step, int arith, int adam_l2,
                               int zero_grad, hsk_stream_t stream) {
    using namespace hsk;
    HSK_REQUIRE(p && m && v && g, "hsk_adamw_dense: null pointer");
    HSK_REQUIRE(n >= 0 && step >= 1, "hsk_adamw_dense: n >= 0 and step >= 1 required (n=%lld step=%lld)", (long long)n,
                (long long)step);
    HSK_REQUIRE(aligned16(p) && aligned16(m) && aligned16(v) && aligned16(g), "hsk_adamw_dense: pointers must be 16-byte aligned");
    HSK_REQUIRE(arith == 0 || arith == 1, "hsk_adamw_dense: arith must be 0 (cuda foreach) or 1 (cpu single-tensor)");
    if (n == 0) return HSK_OK;
    // python-double scalar bookkeeping exactly as torch/optim/adam.py does it, then one rounding to fp32
    AdamConsts c;
    fill_consts(c, lr, beta1, beta2, eps, weight_decay, step);
    const bool l2 = adam_l2 != 0 && weight_decay != 0.0;
    const bool decay = adam_l2 == 0 && weight_decay != 0.0;
    const int threads = 256;
    int64_t want = ((n >> 2) + threads - 1) / threads;
    if (want < 1) want = 1;
    int64_t cap = (int64_t)sm_count() * 16;
    int blocks = (int)(want < cap ? want : cap);
    cudaStream_t s = as_stream(stream);
#define HSK_LAUNCH_ADAM(A, L, D, Z) adamw_dense_kernel<A, L, D, Z><<<blocks, threads, 0, s>>>(p, m, v, g, n, c)
#define HSK_ADAM_Z(A, L, D) \
    if (zero_grad) HSK_LAUNCH_ADAM(A, L, D, true); else HSK_LAUNCH_ADAM(A, L, D, false)
#define HSK_ADAM_LD(A)                         \
    if (l2) { HSK_ADAM_Z(A, true, false); }    \
    else if (decay) { HSK_ADAM_Z(A, false, true); } \
    else { HSK_ADAM_Z(A, false, false); }
    if (arith == 0) { HSK_ADAM_LD(0) } else { HSK_ADAM_LD(1) }
    return check_launch("hsk_adamw_dense");
}

// ---- the SAME dense update with a per-row "touched" stamp: rows nobody scattered a gradient into this step have an
// all-zero gradient row (the optimizer leaves g zero, the scatter kernels are the only writers), so their g is neither
// read nor re-zeroed: 24 B / element (read + write p, m, v) instead of 32 B.  With g = 0 the arithmetic is the arithmetic
// of adamw_dense_kernel on a zero gradient, hence bit-identical results.  At cfg4 (2 M x 1 M x 128, B 8192) a step
// touches 0.35 M of 3 M rows.  A row counts as touched when stamps[row] == stamp; the stamp of step t is
// 1 + t % 255, so the byte array never needs clearing (a stale match after 255 steps only costs that row's g traffic).
namespace hsk {

struct AdamSeg {
    int64_t begin4, end4;            // float4 range [begin4, end4) of the arena
    const uint8_t* stamps;           // one byte per row of the segment
    uint64_t magic;                  // ceil(2^64 / nvec): row = umul64hi(float4 index inside the segment, magic); 0: nvec = 1
};
struct AdamSegs {
    AdamSeg s[4];
    int n;
    int64_t lo4, hi4;                // float4 range covered by the kernel: [lo4, hi4) = first segment's begin .. last one's end
};

__host__ __device__ __forceinline__ int stamp_of_step(int64_t step) { return 1 + (int)(step % 255); }

// Flat streaming pass (one float4 per thread and iteration, grid-stride — the access pattern and occupancy of
// adamw_dense_kernel, which reaches 94 % of the copy bandwidth) over the float4 range spanned by the row segments.  The
// p / m / v loads are issued first; the row of the float4 comes from a multiply-high (no integer division), its stamp byte
// is one more independent load; only then the gradient is loaded — for stamped rows only.  float4s between two segments
// (alignment padding) take the dense path.  Measured alternatives at cfg4 (385 M parameters, 0.35 M of 3 M rows stamped):
// row lookup by integer division with the stamp load in front of the data loads 1.88 ms; warp-per-row layouts with 4 / 2 /
// 1 rows per warp 2.34-2.47 / 2.08 / 1.93 ms (fewer resident warps hide the ~60 instructions per element of IEEE
// division / square root worse); the plain 32 B / parameter kernel 2.01 ms.
template <int ARITH, bool L2, bool DECAY>
__global__ void __launch_bounds__(256) adamw_rows_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                         float* __restrict__ g, AdamConsts c_host,
                                                         const AdamConsts* __restrict__ c_dev, AdamSegs segs, int stamp_host,
                                                         const int64_t* __restrict__ step_dev) {
    const AdamConsts c = c_dev ? *c_dev : c_host;
    const uint8_t stamp = (uint8_t)(step_dev ? stamp_of_step(*step_dev) : stamp_host);
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* g4 = reinterpret_cast<float4*>(g);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = segs.lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < segs.hi4; i += stride) {
        float4 P = p4[i], M = m4[i], V = v4[i];
        uint8_t st = stamp;          // outside every segment: dense path
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < segs.n && i >= segs.s[k].begin4 && i < segs.s[k].end4)
                st = __ldg(segs.s[k].stamps + (segs.s[k].magic ? __umul64hi((uint64_t)(i - segs.s[k].begin4), segs.s[k].magic)
                                                             : (uint64_t)(i - segs.s[k].begin4)));
        }
        const bool touched = st == stamp;
        float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
        if (touched) G = __ldcs(g4 + i);
        adam_elem<ARITH, L2, DECAY>(P.x, M.x, V.x, G.x, c);
        adam_elem<ARITH, L2, DECAY>(P.y, M.y, V.y, G.y, c);
        adam_elem<ARITH, L2, DECAY>(P.z, M.z, V.z, G.z, c);
        adam_elem<ARITH, L2, DECAY>(P.w, M.w, V.w, G.w, c);
        p4[i] = P;
        m4[i] = M;
        v4[i] = V;
        if (touched) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(256) mark_rows_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t n_rows,
                                                        uint8_t* __restrict__ stamps, int stamp_host,
                                                        const int64_t* __restrict__ step_dev) {
    const uint8_t stamp = (uint8_t)(step_dev ? stamp_of_step(*step_dev) : stamp_host);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int64_t r = idx[e];
        if (!bad_index(r, n_rows)) stamps[r] = stamp;
    }
}

__global__ void __launch_bounds__(256) mark_batch_kernel(const int64_t* __restrict__ u_idx, const int64_t* __restrict__ i_idx,
                                                         int64_t B, int64_t BN1, int64_t n_users, int64_t n_items,
                                                         uint8_t* __restrict__ su, uint8_t* __restrict__ si, int stamp_host,
                                                         const int64_t* __restrict__ step_dev) {
    const uint8_t stamp = (uint8_t)(step_dev ? stamp_of_step(*step_dev) : stamp_host);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < BN1; e += stride) {
        if (si) {
            const int64_t it = i_idx[e];
            if (!bad_index(it, n_items)) si[it] = stamp;
        }
        if (su && e < B) {
            const int64_t u = u_idx[e];
            if (!bad_index(u, n_users)) su[u] = stamp;
        }
    }
}

}  // namespace hsk

extern "C" int hsk_row_stamp(int64_t step) { return hsk::stamp_of_step(step); }

extern "C" int hsk_mark_batch(const int64_t* u_idx, const int64_t* i_idx, int B, int N1, int64_t n_users, int64_t n_items,
                              uint8_t* stamps_users, uint8_t* stamps_items, int64_t step, const int64_t* step_dev,
                              hsk_stream_t stream) {
    HSK_REQUIRE(u_idx && i_idx, "hsk_mark_batch: null pointer");
    if (B <= 0 || (!stamps_users && !stamps_items)) return HSK_OK;
    const int64_t n = (int64_t)B * N1;
    int64_t blocks = (n + 255) / 256, cap = (int64_t)hsk::sm_count() * 8;
    hsk::mark_batch_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, hsk::as_stream(stream)>>>(
        u_idx, i_idx, B, n, n_users, n_items, stamps_users, stamps_items, hsk::stamp_of_step(step), step_dev);
    return hsk::check_launch("hsk_mark_batch");
}

extern "C" int hsk_mark_rows(const int64_t* idx, int64_t n, int64_t n_rows, uint8_t* stamps, int64_t step,
                             const int64_t* step_dev, hsk_stream_t stream) {
    HSK_REQUIRE(n >= 0 && n_rows >= 0, "hsk_mark_rows: bad sizes");
    if (n == 0) return HSK_OK;
    HSK_REQUIRE(idx && stamps, "hsk_mark_rows: null pointer");
    int64_t blocks = (n + 255) / 256, cap = (int64_t)hsk::sm_count() * 8;
    hsk::mark_rows_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, hsk::as_stream(stream)>>>(idx, n, n_rows, stamps,
                                                                                            hsk::stamp_of_step(step), step_dev);
    return hsk::check_launch("hsk_mark_rows");
}

extern "C" int hsk_adamw_dense_rows(float* p, float* m, float* v, float* g, int64_t n, const hsk_row_segment* segments,
                                    int n_segments, double lr, double beta1, double beta2, double eps, double weight_decay,
                                    int64_t step, const float* consts_dev, const int64_t* step_dev, int arith, int adam_l2,
                                    hsk_stream_t stream) {
    using namespace hsk;
    HSK_REQUIRE(p && m && v && g, "hsk_adamw_dense_rows: null pointer");
    HSK_REQUIRE(n >= 0 && (consts_dev || step >= 1), "hsk_adamw_dense_rows: n >= 0 and step >= 1 required");
    HSK_REQUIRE(aligned16(p) && aligned16(m) && aligned16(v) && aligned16(g), "hsk_adamw_dense_rows: pointers must be 16-byte aligned");
    HSK_REQUIRE(arith == 0 || arith == 1, "hsk_adamw_dense_rows: arith must be 0 (cuda foreach) or 1 (cpu single-tensor)");
    HSK_REQUIRE(n_segments >= 0 && n_segments <= 4 && (n_segments == 0 || segments), "hsk_adamw_dense_rows: at most 4 row segments");
    HSK_REQUIRE((consts_dev == nullptr) == (step_dev == nullptr), "hsk_adamw_dense_rows: consts_dev and step_dev go together (graph mode)");
    HSK_REQUIRE(!consts_dev || aligned16(consts_dev), "hsk_adamw_dense_rows: consts_dev must be 16-byte aligned");
    HSK_REQUIRE(!consts_dev || arith == 0, "hsk_adamw_dense_rows: graph mode supports arith 0 only");
    if (n == 0) return HSK_OK;
    AdamSegs segs;
    memset(&segs, 0, sizeof(segs));
    int64_t cursor = 0;   // segments must be ascending and disjoint
    for (int k = 0; k < n_segments; ++k) {
        const hsk_row_segment& sg = segments[k];
        HSK_REQUIRE(sg.n_rows >= 0 && sg.ld >= 4 && sg.ld % 4 == 0 && sg.offset % 4 == 0 && sg.offset >= cursor &&
                        sg.offset + sg.n_rows * sg.ld <= n && sg.stamps,
                    "hsk_adamw_dense_rows: segment %d must be a 16-byte aligned range of rows (ld %% 4 == 0) inside [0, n), "
                    "segments ascending and disjoint", k);
        HSK_REQUIRE(sg.n_rows * (int64_t)(sg.ld / 4) < ((int64_t)1 << 32), "hsk_adamw_dense_rows: segment %d too long", k);
        cursor = sg.offset + sg.n_rows * sg.ld;
        if (sg.n_rows == 0) continue;
        const uint64_t nvec = (uint64_t)(sg.ld / 4);
        AdamSeg& d = segs.s[segs.n++];
        d.begin4 = sg.offset / 4;
        d.end4 = (sg.offset + sg.n_rows * sg.ld) / 4;
        d.stamps = sg.stamps;
        d.magic = nvec == 1 ? 0 : (~0ull) / nvec + 1;     // ceil(2^64 / nvec); 0 = one float4 per row (row = index)
    }
    // what lies before the first / after the last segment (bias vectors, the global bias) takes the plain dense kernel; the
    // alignment padding BETWEEN segments is walked by the rows kernel on its dense path
    struct Gap { int64_t off, len; } gaps[2];
    int n_gaps = 0;
    if (segs.n > 0) {
        segs.lo4 = segs.s[0].begin4;
        segs.hi4 = segs.s[segs.n - 1].end4;
        if (segs.lo4 > 0) gaps[n_gaps++] = {0, segs.lo4 * 4};
        if (segs.hi4 * 4 < n) gaps[n_gaps++] = {segs.hi4 * 4, n - segs.hi4 * 4};
    } else {
        gaps[n_gaps++] = {0, n};
    }
    AdamConsts c;
    memset(&c, 0, sizeof(c));
    if (!consts_dev) fill_consts(c, lr, beta1, beta2, eps, weight_decay, step);
    // graph mode: the decay scalars come from the device table (a zero weight decay multiplies by exactly 1 / adds exactly 0)
    const bool l2 = adam_l2 != 0 && (consts_dev || weight_decay != 0.0);
    const bool decay = adam_l2 == 0 && (consts_dev || weight_decay != 0.0);
    cudaStream_t s = as_stream(stream);
    const AdamConsts* cp = reinterpret_cast<const AdamConsts*>(consts_dev);
    const int stamp = stamp_of_step(step);
    if (segs.n > 0) {
        const int64_t want = (segs.hi4 - segs.lo4 + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        const int blocks = (int)(want < cap ? want : cap);
#define HSK_LAUNCH_ADAMR(A, L, D) adamw_rows_kernel<A, L, D><<<blocks, 256, 0, s>>>(p, m, v, g, c, cp, segs, stamp, step_dev)
#define HSK_ADAMR_LD(A)                                 \
    if (l2) { HSK_LAUNCH_ADAMR(A, true, false); }       \
    else if (decay) { HSK_LAUNCH_ADAMR(A, false, true); } \
    else { HSK_LAUNCH_ADAMR(A, false, false); }
        if (arith == 0) { HSK_ADAMR_LD(0) } else { HSK_ADAMR_LD(1) }
        int rc = check_launch("hsk_adamw_dense_rows");
        if (rc) return rc;
    }
    // bias vectors, the global bias, alignment padding: the plain streaming kernel on what the segments leave out
    for (int k = 0; k < n_gaps; ++k) {
        int rc;
        if (consts_dev)
            rc = hsk_adamw_dense_graph(p + gaps[k].off, m + gaps[k].off, v + gaps[k].off, g + gaps[k].off, gaps[k].len, consts_dev,
                                       adam_l2 == 0, adam_l2, 1, stream);
        else
            rc = hsk_adamw_dense(p + gaps[k].off, m + gaps[k].off, v + gaps[k].off, g + gaps[k].off, gaps[k].len, lr, beta1, beta2,
                                 eps, weight_decay, step, arith, adam_l2, 1, stream);
        if (rc) return rc;
    }
    return HSK_OK;
}

// ---- row-sparse "lazy" AdamW (north_star item 2, reported separately from the dense, torch-faithful mode) ------------
// Only rows that received a gradient in this step are updated (p, m, v) — the semantics of torch.optim.SparseAdam plus
// decoupled weight decay on the touched rows; untouched rows keep p, m, v unchanged (no decay, no momentum tail), which
// is a different trajectory from dense AdamW.  Bytes: 28 B per TOUCHED element + 1 flag byte per row,
// instead of 28 B per parameter: at cfg4 (2 M x 1 M x 128) a step touches ~0.35 M of 3 M rows.
namespace hsk {

__global__ void __launch_bounds__(256) mark_touched_kernel(const int64_t* __restrict__ u_idx, const int64_t* __restrict__ i_idx,
                                                           int64_t B, int64_t BN1, int64_t n_users, int64_t n_items,
                                                           uint8_t* __restrict__ tu, uint8_t* __restrict__ ti) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < BN1; e += stride) {
        const int64_t it = i_idx[e];
        if (!bad_index(it, n_items)) ti[it] = 1;
        if (e < B) {
            const int64_t u = u_idx[e];
            if (!bad_index(u, n_users)) tu[u] = 1;
        }
    }
}

// one warp per row; bias (nullable) is the length-n_rows vector that shares the row index (item_bias / user_bias)
template <bool DECAY>
__global__ void __launch_bounds__(256) adamw_rows_lazy_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                              float* __restrict__ g, int64_t n_rows, int ld,
                                                              float* __restrict__ pb, float* __restrict__ mb, float* __restrict__ vb,
                                                              float* __restrict__ gb, uint8_t* __restrict__ touched, AdamConsts c) {
    const int lane = threadIdx.x & 31;
    const int nvec = ld >> 2;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_rows; r += (int64_t)gridDim.x * 8) {
        if (!touched[r]) continue;
        float4* p4 = reinterpret_cast<float4*>(p + r * ld);
        float4* m4 = reinterpret_cast<float4*>(m + r * ld);
        float4* v4 = reinterpret_cast<float4*>(v + r * ld);
        float4* g4 = reinterpret_cast<float4*>(g + r * ld);
        for (int k = lane; k < nvec; k += 32) {
            float4 P = p4[k], M = m4[k], V = v4[k], G = g4[k];
            adam_elem<0, false, DECAY>(P.x, M.x, V.x, G.x, c);
            adam_elem<0, false, DECAY>(P.y, M.y, V.y, G.y, c);
            adam_elem<0, false, DECAY>(P.z, M.z, V.z, G.z, c);
            adam_elem<0, false, DECAY>(P.w, M.w, V.w, G.w, c);
            p4[k] = P; m4[k] = M; v4[k] = V;
            g4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (pb && lane == 0) {
            float P = pb[r], M = mb[r], V = vb[r], G = gb[r];
            adam_elem<0, false, DECAY>(P, M, V, G, c);
            pb[r] = P; mb[r] = M; vb[r] = V; gb[r] = 0.f;
        }
        __syncwarp();
        if (lane == 0) touched[r] = 0;
    }
}

}  // namespace hsk

extern "C" int hsk_mark_touched(const int64_t* u_idx, const int64_t* i_idx, int B, int N1, int64_t n_users, int64_t n_items,
                                uint8_t* touched_users, uint8_t* touched_items, hsk_stream_t stream) {
    HSK_REQUIRE(u_idx && i_idx && touched_users && touched_items, "hsk_mark_touched: null pointer");
    if (B <= 0) return HSK_OK;
    const int64_t n = (int64_t)B * N1;
    int64_t blocks = (n + 255) / 256, cap = (int64_t)hsk::sm_count() * 8;
    hsk::mark_touched_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, hsk::as_stream(stream)>>>(u_idx, i_idx, B, n, n_users, n_items,
                                                                                                 touched_users, touched_items);
    return hsk::check_launch("hsk_mark_touched");
}

extern "C" int hsk_adamw_rows_lazy(float* p, float* m, float* v, float* g, int64_t n_rows, int ld, float* p_bias, float* m_bias,
                                   float* v_bias, float* g_bias, uint8_t* touched, double lr, double beta1, double beta2,
                                   double eps, double weight_decay, int64_t step, hsk_stream_t stream) {
    using namespace hsk;
    HSK_REQUIRE(p && m && v && g && touched, "hsk_adamw_rows_lazy: null pointer");
    HSK_REQUIRE(ld >= 4 && (ld % 4) == 0 && aligned16(p) && aligned16(m) && aligned16(v) && aligned16(g), "hsk_adamw_rows_lazy: rows must be 16-byte aligned with ld %% 4 == 0");
    HSK_REQUIRE(step >= 1 && n_rows >= 0, "hsk_adamw_rows_lazy: bad step / row count");
    HSK_REQUIRE((p_bias == nullptr) == (m_bias == nullptr) && (p_bias == nullptr) == (v_bias == nullptr) && (p_bias == nullptr) == (g_bias == nullptr),
                "hsk_adamw_rows_lazy: bias p/m/v/g must be given together");
    if (n_rows == 0) return HSK_OK;
    AdamConsts c;
    fill_consts(c, lr, beta1, beta2, eps, weight_decay, step);
    int64_t blocks = (n_rows + 7) / 8, cap = (int64_t)sm_count() * 16;
    const int nb = (int)(blocks < cap ? blocks : cap);
    if (weight_decay != 0.0)
        adamw_rows_lazy_kernel<true><<<nb, 256, 0, as_stream(stream)>>>(p, m, v, g, n_rows, ld, p_bias, m_bias, v_bias, g_bias, touched, c);
    else
        adamw_rows_lazy_kernel<false><<<nb, 256, 0, as_stream(stream)>>>(p, m, v, g, n_rows, ld, p_bias, m_bias, v_bias, g_bias, touched, c);
    return check_launch("hsk_adamw_rows_lazy");
}

// ---- torch.optim.Adagrad(params, lr, weight_decay) (train/trainer.py:50-51), dense, one streaming pass ----------------
// Op order of torch's CUDA `_multi_tensor_adagrad` (lr_decay = 0, initial_accumulator_value = 0, eps = 1e-10):
//   g = fma(f(wd), p, g)            _foreach_add(grads, params, alpha = wd)          (only if wd != 0)
//   s = fma(g, g, s)                _foreach_addcmul_(state_sums, grads, grads, 1)   (value 1 is folded away)
//   std = sqrt(s) + f(eps)          _foreach_sqrt, _foreach_add_
//   p = p + (g * f(-lr)) / std      _foreach_mul_(grads, -clr), _foreach_addcdiv_(params, grads, std)
namespace hsk {
template <bool L2, bool ZERO>
__global__ void __launch_bounds__(256) adagrad_dense_kernel(float* __restrict__ p, float* __restrict__ s, float* __restrict__ g,
                                                            int64_t n, float wd, float neg_lr, float eps) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float P = p[i], S = s[i], G = g[i];
        if (L2) G = __fmaf_rn(wd, P, G);
        S = __fmaf_rn(G, G, S);   // value = 1 is folded: a single rounding (verified bitwise against torch on the GPU)
        const float sd = __fadd_rn(__fsqrt_rn(S), eps);
        P = __fmaf_rn(1.0f, __fdiv_rn(__fmul_rn(G, neg_lr), sd), P);
        p[i] = P; s[i] = S;
        if (ZERO) g[i] = 0.f;
    }
}
}  // namespace hsk

extern "C" int hsk_adagrad_dense(float* p, float* state_sum, float* g, int64_t n, double lr, double eps, double weight_decay,
                                 int zero_grad, hsk_stream_t stream) {
    using namespace hsk;
    HSK_REQUIRE(p && state_sum && g && n >= 0, "hsk_adagrad_dense: bad arguments");
    if (n == 0) return HSK_OK;
    int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count() * 32;
    const int nb = (int)(blocks < cap ? blocks : cap);
    cudaStream_t st = as_stream(stream);
    const float wd = (float)weight_decay, nlr = (float)(-lr), e = (float)eps;
    if (weight_decay != 0.0) {
        if (zero_grad) adagrad_dense_kernel<true, true><<<nb, 256, 0, st>>>(p, state_sum, g, n, wd, nlr, e);
        else adagrad_dense_kernel<true, false><<<nb, 256, 0, st>>>(p, state_sum, g, n, wd, nlr, e);
    } else {
        if (zero_grad) adagrad_dense_kernel<false, true><<<nb, 256, 0, st>>>(p, state_sum, g, n, wd, nlr, e);
        else adagrad_dense_kernel<false, false><<<nb, 256, 0, st>>>(p, state_sum, g, n, wd, nlr, e);
    }
    return check_launch("hsk_adagrad_dense");
}
