// Peer mapping of device allocations between the ranks of one node (CUDA IPC) for the peer exchange of the item-sharded
// step (hsk_mf_train_fused_peer): a rank exports the allocation behind a tensor, its peers map it once and address it like
// local memory (loads, stores and atomics travel over NVLink / NVSwitch).
#include <cuda.h>

#include "hsk_common.cuh"

namespace hsk {

typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

static int address_range(const void* ptr, CUdeviceptr* base, size_t* size) {
    static GetAddressRangeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            return set_err(HSK_ERR_CUDA, "cuMemGetAddressRange entry point unavailable: %s", cudaGetErrorString(e));
        fn = reinterpret_cast<GetAddressRangeFn>(p);
    }
    CUresult r = fn(base, size, reinterpret_cast<CUdeviceptr>(ptr));
    if (r != CUDA_SUCCESS) return set_err(HSK_ERR_CUDA, "cuMemGetAddressRange failed (%d)", (int)r);
    return HSK_OK;
}

}  // namespace hsk

namespace hsk {

constexpr long long kBarrierSpinCycles = 40ll * 1000 * 1000 * 1000;   // ~20 s at 2 GHz

__global__ void __launch_bounds__(32) peer_barrier_kernel(hsk_peer_flags f, uint32_t* epoch, int32_t* status) {
    __shared__ uint32_t e_s;
    const int q = threadIdx.x;
    if (q == 0) e_s = *epoch + 1u;
    __syncthreads();
    const uint32_t e = e_s;
    if (q < f.world) {
        // everything this GPU did before (previous kernels on the stream, their peer stores / reductions included) is
        // ordered before the arrival flag
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(f.flags[q] + f.rank), "r"(e) : "memory");
        const uint32_t* mine = f.flags[f.rank] + q;
        const long long t0 = clock64();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(mine) : "memory");
            if ((int32_t)(v - e) >= 0) break;          // rank q has arrived at this barrier (or already at the next one)
            if (clock64() - t0 > kBarrierSpinCycles) {
                if (status) atomicOr(status, HSK_STATUS_BARRIER_TIMEOUT);
                break;
            }
        }
    }
    __syncthreads();
    if (q == 0) *epoch = e;
}

}  // namespace hsk

extern "C" int hsk_peer_barrier(const hsk_peer_flags* f, uint32_t* epoch, int32_t* status, hsk_stream_t stream) {
    HSK_REQUIRE(f && epoch, "hsk_peer_barrier: null pointer");
    HSK_REQUIRE(f->world >= 1 && f->world <= HSK_MAX_PEERS && f->rank >= 0 && f->rank < f->world, "hsk_peer_barrier: bad world / rank");
    for (int q = 0; q < f->world; ++q) HSK_REQUIRE(f->flags[q], "hsk_peer_barrier: flags of rank %d missing", q);
    hsk::peer_barrier_kernel<<<1, 32, 0, hsk::as_stream(stream)>>>(*f, epoch, status);
    return hsk::check_launch("hsk_peer_barrier");
}

static_assert(sizeof(cudaIpcMemHandle_t) == HSK_PEER_HANDLE_BYTES, "IPC handle size");

extern "C" int hsk_peer_export(const void* ptr, void* handle, int64_t* offset) {
    HSK_REQUIRE(ptr && handle && offset, "hsk_peer_export: null pointer");
    CUdeviceptr base = 0;
    size_t size = 0;
    int rc = hsk::address_range(ptr, &base, &size);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return hsk::set_err(HSK_ERR_CUDA, "hsk_peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
    *offset = (int64_t)(reinterpret_cast<CUdeviceptr>(ptr) - base);
    return HSK_OK;
}

extern "C" int hsk_peer_open(const void* handle, void** base_out) {
    HSK_REQUIRE(handle && base_out, "hsk_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return hsk::set_err(HSK_ERR_CUDA, "hsk_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    }
    *base_out = p;
    return HSK_OK;
}

extern "C" int hsk_peer_close(void* base) {
    if (!base) return HSK_OK;
    cudaError_t e = cudaIpcCloseMemHandle(base);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return hsk::set_err(HSK_ERR_CUDA, "hsk_peer_close: cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
    }
    return HSK_OK;
}
