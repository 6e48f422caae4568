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
