// Shared device/host helpers for the hassaku_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "hassaku_b200.h"

namespace hsk {

// ---- host: thread-local error text (hsk_last_error) ----
char* err_buf();
int set_err(int code, const char* fmt, ...);
int check_launch(const char* what);

#define HSK_REQUIRE(cond, ...)                                        \
    do {                                                              \
        if (!(cond)) return hsk::set_err(HSK_ERR_INVALID, __VA_ARGS__); \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline cudaStream_t as_stream(hsk_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// 128-bit read-only load of a table row fragment (rows are re-read by other CTAs: keep them in L1/L2)
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// A gathered embedding row held across a warp: lane l owns float4 #(l + 32 k), k < NV.
template <int NV>
struct Row {
    float4 v[NV];
    __device__ __forceinline__ void load(const float* __restrict__ base, int nvec, int lane) {
        const float4* p = reinterpret_cast<const float4*>(base);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            int idx = lane + 32 * k;
            v[k] = (idx < nvec) ? ldg4(p + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ float dot_partial(const Row<NV>& o) const {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            a = fmaf(v[k].x, o.v[k].x, a);
            a = fmaf(v[k].y, o.v[k].y, a);
            a = fmaf(v[k].z, o.v[k].z, a);
            a = fmaf(v[k].w, o.v[k].w, a);
        }
        return a;
    }
    __device__ __forceinline__ void axpy(float a, const Row<NV>& x) {  // this += a * x
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            v[k].x = fmaf(a, x.v[k].x, v[k].x);
            v[k].y = fmaf(a, x.v[k].y, v[k].y);
            v[k].z = fmaf(a, x.v[k].z, v[k].z);
            v[k].w = fmaf(a, x.v[k].w, v[k].w);
        }
    }
    // dst[row] += a * this  — 128-bit vector reductions into global memory (RED.E.ADD.F32x4 on sm_90+)
    __device__ __forceinline__ void red_scaled(float* __restrict__ dst, float a, int nvec, int lane) const {
        float4* p = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            int idx = lane + 32 * k;
            if (idx < nvec) atomicAdd(p + idx, make_float4(a * v[k].x, a * v[k].y, a * v[k].z, a * v[k].w));
        }
    }
    __device__ __forceinline__ void red(float* __restrict__ dst, int nvec, int lane) const {
        float4* p = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            int idx = lane + 32 * k;
            if (idx < nvec) atomicAdd(p + idx, v[k]);
        }
    }
};

__device__ __forceinline__ bool bad_index(int64_t idx, int64_t n) { return static_cast<uint64_t>(idx) >= static_cast<uint64_t>(n); }

// fp32 log-sigmoid with torch's formulation: min(0, x) - log1p(exp(-|x|))
__device__ __forceinline__ float log_sigmoid_f(float x) { return fminf(0.f, x) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// ---- mbarrier + 1-D bulk async copy (TMA engine, UBLKCP): global -> shared, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    asm volatile("fence.proxy.async.shared::cta;\n" ::);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes % 16 == 0, src and dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// the same on precomputed 32-bit shared-memory addresses (hot loops: no generic->shared conversion per call)
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src_gmem),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
    return v;
}

}  // namespace hsk

// dispatch on the number of float4 per lane: NV = ceil((ld / 4) / 32), ld <= 1024
#define HSK_DISPATCH_NV(nv, ...)                                             \
    switch (nv) {                                                            \
        case 1: { constexpr int NV = 1; __VA_ARGS__; } break;                \
        case 2: { constexpr int NV = 2; __VA_ARGS__; } break;                \
        case 3: { constexpr int NV = 3; __VA_ARGS__; } break;                \
        case 4: { constexpr int NV = 4; __VA_ARGS__; } break;                \
        case 5: { constexpr int NV = 5; __VA_ARGS__; } break;                \
        case 6: { constexpr int NV = 6; __VA_ARGS__; } break;                \
        case 7: { constexpr int NV = 7; __VA_ARGS__; } break;                \
        case 8: { constexpr int NV = 8; __VA_ARGS__; } break;                \
        default: return hsk::set_err(HSK_ERR_UNSUPPORTED, "embedding_dim > 1024 is not supported (ld=%d)", (int)(nv) * 128); \
    }
