// C-ABI plumbing: thread-local error text, version, device info.
#include <stdarg.h>

#include "hsk_common.cuh"

namespace hsk {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(HSK_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return HSK_OK;
}

int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

}  // namespace hsk

extern "C" {

const char* hsk_last_error(void) { return hsk::err_buf(); }

int hsk_version(void) { return 100; }

int hsk_device_info(int* sm, int* cc_major, int* cc_minor, int64_t* l2_bytes, int64_t* hbm_bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return hsk::set_err(HSK_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return hsk::set_err(HSK_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (sm) *sm = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (l2_bytes) *l2_bytes = p.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = (int64_t)p.totalGlobalMem;
    return HSK_OK;
}

}  // extern "C"
