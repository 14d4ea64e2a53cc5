// Error plumbing + device queries for the C ABI.
#include "common.cuh"
#include <stdarg.h>
#include <mutex>
#include <set>
#include <utility>

namespace vd {

static thread_local char g_err[512] = "";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::mutex g_cfg_mutex;      // guards the per-device caches below (one process may drive several devices / threads)

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// Function attributes are per (device, function): set them once for each pair (not once per process), under a lock.
int configure_kernel(const void* func, int dyn_smem_bytes, bool max_shared_carveout) {
    int dev = 0;
    VD_CUDA(cudaGetDevice(&dev));
    static std::set<std::pair<int, const void*>> done;
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    if (done.count(std::make_pair(dev, func))) return VD_OK;
    if (dyn_smem_bytes > 0) VD_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem_bytes));
    if (max_shared_carveout) VD_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    done.insert(std::make_pair(dev, func));
    return VD_OK;
}

}  // namespace vd

extern "C" int vd_version(void) { return 100; }

// ---- peer-mapped buffers (CUDA IPC) for the fused head's output mirrors
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
extern "C" int vd_ipc_alloc(size_t bytes, void** dev_ptr_out, unsigned char handle_out[64]) {
    VD_CHECK_ARG(bytes > 0 && dev_ptr_out && handle_out, "ipc_alloc: bad argument");
    void* p = nullptr;
    VD_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return vd::set_error(VD_ERR_CUDA, "ipc_alloc: %s", cudaGetErrorString(e)); }
    memcpy(handle_out, &h, 64);
    *dev_ptr_out = p;
    return VD_OK;
}
extern "C" int vd_ipc_open(const unsigned char handle[64], void** dev_ptr_out) {
    VD_CHECK_ARG(handle && dev_ptr_out, "ipc_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    VD_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return VD_OK;
}
extern "C" int vd_ipc_close(void* dev_ptr) { if (dev_ptr) VD_CUDA(cudaIpcCloseMemHandle(dev_ptr)); return VD_OK; }
extern "C" int vd_ipc_free(void* dev_ptr) { if (dev_ptr) VD_CUDA(cudaFree(dev_ptr)); return VD_OK; }
extern "C" const char* vd_last_error(void) { return vd::last_error_buf(); }

extern "C" int vd_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    int n = 0, maj = 0, min = 0;
    VD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    VD_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, device));
    VD_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, device));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (maj != 10)
        return vd::set_error(VD_ERR_UNSUPPORTED, "device %d is sm_%d%d; viddet_b200 is built for sm_100a only", device, maj, min);
    return VD_OK;
}
