// Shared host/device helpers for the viddet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/viddet_b200.h"

namespace vd {

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message, never throws across the C ABI)
// ---------------------------------------------------------------------------------------------
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define VD_CHECK_ARG(cond, ...)                                            \
    do { if (!(cond)) return vd::set_error(VD_ERR_INVALID_ARG, __VA_ARGS__); } while (0)
#define VD_CUDA(call)                                                       \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess)                    \
        return vd::set_error(VD_ERR_CUDA, "%s failed: %s (%s:%d)", #call,   \
                             cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define VD_LAUNCH_CHECK()  VD_CUDA(cudaGetLastError())

// Device-side index checks of the global-memory stores / gathers of the hot kernels.  compute-sanitizer is closed on this pool
// (it left GPUs needing a reset), so the memory-safety evidence is a build variant with these checks compiled in
// (`python -m viddet_b200.build --variant bounds -DVD_BOUNDS_CHECK`, VD_LIB=... pytest -m gpu; profiles/r02_bounds_check.txt): a
// violated check prints its location and traps, which fails the test with a CUDA error.  Compiled out of the product build.
#ifdef VD_BOUNDS_CHECK
#define VD_DEV_CHECK(cond)                                                                                          \
    do { if (!(cond)) { printf("viddet_b200 bounds check failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, \
                               (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define VD_DEV_CHECK(cond) do { } while (0)
#endif

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();   // of the current device (cached per device)
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize[, carve-out]) once per (current device, kernel); thread-safe
int configure_kernel(const void* func, int dyn_smem_bytes, bool max_shared_carveout);
// vd_temporal_conv_ex with a device-side condition: cond != nullptr -> the kernels return at once when *cond == 0; with frame_list the
// CTA-pair kernel computes only the windows of frames frame_list[0 .. *cond)
int temporal_conv_impl(const void* x, void* y, int B, int T, int H, int W, int C, const void* weight, const float* scale, const float* shift,
                       float slope, int precision, int window_stride_frames, const unsigned int* cond, const unsigned int* frame_list, void* stream);

// ---------------------------------------------------------------------------------------------
// 64-bit selection key: (orderable(score) << 32) | ~row.  Larger key == earlier in MXNet's
// stable descending sort (score desc, original row asc).  0 is reserved for "invalid".
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t orderable_f32(float s) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(__fadd_rn(s, 0.0f));      // -0.0 -> +0.0 (ties by row, not by sign)
#else
    float t = s + 0.0f; uint32_t b; memcpy(&b, &t, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float unorderable_f32(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return ((uint64_t)orderable_f32(score) << 32) | (uint32_t)(~row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~(uint32_t)k; }
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return unorderable_f32((uint32_t)(k >> 32)); }

// ---------------------------------------------------------------------------------------------
// decode math shared by the materialising decode kernel and the fused head epilogue, so that the
// two paths produce bit-identical scores and boxes from the same logits (yolo3.py:172-177).
// ---------------------------------------------------------------------------------------------
// ex2 / rcp in their .ftz approx forms: 4 instructions per sigmoid (FMUL, MUFU.EX2, FADD, MUFU.RCP);
// relative error ~1e-6 for |x| < 20, far inside the 1e-5 parity bar against the fp32 oracle.
__device__ __forceinline__ float vd_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float vd_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float vd_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float vd_exp(float x) { return vd_ex2(__fmul_rn(x, 1.4426950408889634f)); }
__device__ __forceinline__ float vd_sigmoid(float x) {
    return vd_rcp(__fadd_rn(1.0f, vd_ex2(__fmul_rn(x, -1.4426950408889634f))));
}
__device__ __forceinline__ float vd_score(float cls_logit, float conf) {
    return __fmul_rn(vd_sigmoid(cls_logit), conf);
}
struct Box4 { float x1, y1, x2, y2; };
__device__ __forceinline__ Box4 vd_decode_box(float tx, float ty, float tw, float th, float gx,
                                              float gy, float stride, float aw, float ah) {
    float cx = __fmul_rn(__fadd_rn(vd_sigmoid(tx), gx), stride);
    float cy = __fmul_rn(__fadd_rn(vd_sigmoid(ty), gy), stride);
    float hw = __fmul_rn(__fmul_rn(vd_exp(tw), aw), 0.5f);   // (exp(tw)*aw)/2.0
    float hh = __fmul_rn(__fmul_rn(vd_exp(th), ah), 0.5f);
    Box4 b;
    b.x1 = __fsub_rn(cx, hw); b.y1 = __fsub_rn(cy, hh);
    b.x2 = __fadd_rn(cx, hw); b.y2 = __fadd_rn(cy, hh);
    return b;
}

// ---------------------------------------------------------------------------------------------
// geometry of a multi-scale head: rows of the concatenated (frames, rows, 6) tensor
// ---------------------------------------------------------------------------------------------
struct HeadGeom {
    int num_scales, num_class, A;
    int H[VD_MAX_SCALES], W[VD_MAX_SCALES], HW[VD_MAX_SCALES];
    int row_base[VD_MAX_SCALES + 1];     // first row of scale s in the concatenated det tensor
    int anc_base[VD_MAX_SCALES + 1];     // first anchor slot of scale s (A*sum HW before s)
    float stride[VD_MAX_SCALES];
    float anchors[VD_MAX_SCALES][6];
};

}  // namespace vd
