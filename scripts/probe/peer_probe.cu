// Measurement only (not part of the product library): how fast can an SM kernel READ a peer's memory over NVLink?
//   probe_linear      coalesced float4 streaming read
//   probe_rows_ld     random 512-byte rows, 8 lanes per row, UNROLL rows in flight per group (the quarter-warp step kernel's pattern)
//   probe_rows_bulk   random 512-byte rows by cp.async.bulk into a per-warp ring of STAGES slots (the ring step kernel's pattern)
// Built by scripts/probe/build.sh into gpurun_exp_probe/libpeer_probe.so; driven by scripts/probe/peer_probe.py.
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) probe_linear(const float4* __restrict__ src, int64_t n4, float* out) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "l"(src + i + k * stride));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].w;
    }
    if (acc == 1.2345f) out[0] = acc;
}

template <int UNROLL>
__global__ void __launch_bounds__(64) probe_rows_ld(const float* __restrict__ src, const int32_t* __restrict__ rows, int n_rows, float* out) {
    const int l = threadIdx.x & 7;
    const int group = (blockIdx.x * 64 + threadIdx.x) >> 3, n_groups = (gridDim.x * 64) >> 3;
    float acc = 0.f;
    for (int r0 = group * UNROLL; r0 < n_rows; r0 += n_groups * UNROLL) {
        float4 v[UNROLL][4];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            const int rr = r0 + q < n_rows ? rows[r0 + q] : 0;
            const float4* p = reinterpret_cast<const float4*>(src + (int64_t)rr * 128) + l;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[q][k].x), "=f"(v[q][k].y), "=f"(v[q][k].z), "=f"(v[q][k].w) : "l"(p + 8 * k));
        }
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += v[q][k].x + v[q][k].w;
    }
    if (acc == 1.2345f) out[0] = acc;
}

template <int STAGES>
__global__ void __launch_bounds__(128) probe_rows_bulk(const float* __restrict__ src, const int32_t* __restrict__ rows, int n_rows, float* out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ uint64_t bars[4][STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* ring = dyn + (size_t)warp * STAGES * 512;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[warp][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int gw = blockIdx.x * 4 + warp, n_w = gridDim.x * 4;
    // this warp's rows: gw, gw + n_w, ...
    const int n_my = gw < n_rows ? (n_rows - gw + n_w - 1) / n_w : 0;
    auto issue = [&](int t) {
        if (lane == 0) {
            const int s = t % STAGES;
            const uint32_t bar = smem_u32(&bars[warp][s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(512u) : "memory");
            const float* g = src + (int64_t)rows[gw + t * n_w] * 128;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + s * 512)),
                         "l"(g), "r"(512u), "r"(bar)
                         : "memory");
        }
    };
    for (int t = 0; t < STAGES - 1 && t < n_my; ++t) issue(t);
    float acc = 0.f;
    for (int t = 0; t < n_my; ++t) {
        if (t + STAGES - 1 < n_my) issue(t + STAGES - 1);
        const int s = t % STAGES;
        const uint32_t bar = smem_u32(&bars[warp][s]), parity = (t / STAGES) & 1;
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        acc += reinterpret_cast<const float4*>(ring + s * 512)[lane].x;
        __syncwarp();
    }
    if (acc == 1.2345f) out[0] = acc;
}

extern "C" {
int probe_launch_linear(const void* src, int64_t n_bytes, int blocks, float* out, void* stream) {
    probe_linear<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)src, n_bytes / 16, out);
    return (int)cudaGetLastError();
}
int probe_launch_rows_ld(const void* src, const int32_t* rows, int n_rows, int unroll, int blocks, float* out, void* stream) {
    if (unroll == 2) probe_rows_ld<2><<<blocks, 64, 0, (cudaStream_t)stream>>>((const float*)src, rows, n_rows, out);
    else if (unroll == 4) probe_rows_ld<4><<<blocks, 64, 0, (cudaStream_t)stream>>>((const float*)src, rows, n_rows, out);
    else probe_rows_ld<8><<<blocks, 64, 0, (cudaStream_t)stream>>>((const float*)src, rows, n_rows, out);
    return (int)cudaGetLastError();
}
int probe_launch_rows_bulk(const void* src, const int32_t* rows, int n_rows, int stages, int blocks, float* out, void* stream) {
    if (stages == 4) probe_rows_bulk<4><<<blocks, 128, 4 * 4 * 512, (cudaStream_t)stream>>>((const float*)src, rows, n_rows, out);
    else if (stages == 8) probe_rows_bulk<8><<<blocks, 128, 4 * 8 * 512, (cudaStream_t)stream>>>((const float*)src, rows, n_rows, out);
    else {
        cudaFuncSetAttribute(probe_rows_bulk<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16 * 512);
        probe_rows_bulk<16><<<blocks, 128, 4 * 16 * 512, (cudaStream_t)stream>>>((const float*)src, rows, n_rows, out);
    }
    return (int)cudaGetLastError();
}
}
