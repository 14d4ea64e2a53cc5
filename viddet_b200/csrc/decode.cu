// vd_yolo_decode -- YOLOOutputV3.hybrid_forward after the prediction conv
// (yolo3.py:158-199; twin yolo3_temporal.py:139-179), on a materialised NCHW fp32 `pred`.
// Compatibility surface (the fused head never materialises these rows); HBM-bound elementwise.
//
// One thread per (cell, anchor): reads its 5+C logits (coalesced across cells: pred is NCHW so
// consecutive threads of a channel read consecutive cells), computes box / conf once, then
// writes C detection rows.  Row order: row = c*(HW*A) + cell*A + a.
#include "common.cuh"

namespace vd {

struct DecodeArgs {
    const float* pred; int B, H, W, C, A; float stride; float anchors[12]; int mode;
    float* det; int64_t det_rows_total, det_row_offset;
    float* raw_centers; float* raw_scales; float* objness; float* class_pred;
};

__global__ void __launch_bounds__(256)
yolo_decode_kernel(DecodeArgs a) {
    const int HW = a.H * a.W, P = 5 + a.C;
    const int b = blockIdx.y;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= HW) return;
    const int gx = cell % a.W, gy = cell / a.W;
    const float* pb = a.pred + (size_t)b * a.A * P * HW + cell;
    for (int an = 0; an < a.A; ++an) {
        const float* pa = pb + (size_t)an * P * HW;
        float tx = __ldg(pa), ty = __ldg(pa + HW), tw = __ldg(pa + 2 * (size_t)HW), th = __ldg(pa + 3 * (size_t)HW);
        float to = __ldg(pa + 4 * (size_t)HW);
        Box4 bx = vd_decode_box(tx, ty, tw, th, (float)gx, (float)gy, a.stride, a.anchors[2 * an], a.anchors[2 * an + 1]);
        const int slot = cell * a.A + an;                       // row inside (HW*A)
        if (a.mode == VD_MODE_TRAIN) {
            float* d = a.det + ((size_t)b * a.det_rows_total + a.det_row_offset + slot) * 4;
            d[0] = bx.x1; d[1] = bx.y1; d[2] = bx.x2; d[3] = bx.y2;
            size_t o = (size_t)b * HW * a.A + slot;
            a.raw_centers[o * 2] = tx; a.raw_centers[o * 2 + 1] = ty;
            a.raw_scales[o * 2] = tw; a.raw_scales[o * 2 + 1] = th;
            a.objness[o] = to;
            for (int c = 0; c < a.C; ++c) a.class_pred[o * a.C + c] = __ldg(pa + (size_t)(5 + c) * HW);
            continue;
        }
        float conf = vd_sigmoid(to);
        if (a.mode == VD_MODE_AGNOSTIC) {
            float* d = a.det + ((size_t)b * a.det_rows_total + a.det_row_offset + slot) * 6;
            d[0] = __fadd_rn(__fmul_rn(conf, 0.0f), 0.0f); d[1] = conf;
            d[2] = bx.x1; d[3] = bx.y1; d[4] = bx.x2; d[5] = bx.y2;
            continue;
        }
        for (int c = 0; c < a.C; ++c) {
            float s = vd_score(__ldg(pa + (size_t)(5 + c) * HW), conf);
            size_t row = (size_t)c * HW * a.A + slot;
            float* d = a.det + ((size_t)b * a.det_rows_total + a.det_row_offset + row) * 6;
            // ids = scores*0 + c (yolo3.py:194): NaN/Inf scores give NaN ids
            float2* d2 = reinterpret_cast<float2*>(d);
            d2[0] = make_float2(__fadd_rn(__fmul_rn(s, 0.0f), (float)c), s);
            d2[1] = make_float2(bx.x1, bx.y1);
            d2[2] = make_float2(bx.x2, bx.y2);
        }
    }
}

// (B, C, H, W) fp32 -> (B, H, W, C) bf16 through a 32x32 shared-memory transpose tile.
__global__ void __launch_bounds__(256)
repack_nchw_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const float* s = src + (size_t)b * C * HW;
    __nv_bfloat16* d = dst + (size_t)b * HW * C;
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, p = p0 + tx;
        tile[i][tx] = (c < C && p < HW) ? s[(size_t)c * HW + p] : 0.0f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int p = p0 + i, c = c0 + tx;
        if (p < HW && c < C) d[(size_t)p * C + c] = __float2bfloat16_rn(tile[tx][i]);
    }
}

// fp32 -> bf16 planes: p0 = bf16(v), p1 = bf16(v - p0), p2 = bf16(v - p0 - p1); the subtractions are exact in fp32
__device__ __forceinline__ void split_planes(float v, int planes, __nv_bfloat16* out /*[3]*/) {
    out[0] = __float2bfloat16_rn(v);
    const float r1 = __fsub_rn(v, __bfloat162float(out[0]));
    out[1] = __float2bfloat16_rn(r1);
    out[2] = planes > 2 ? __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(out[1]))) : __float2bfloat16_rn(0.0f);
}
// (B, C, H, W) fp32 -> (planes, B, H, W, C) bf16 planes (carrier of the fp32-parity modes)
__global__ void __launch_bounds__(256)
repack_nchw_to_nhwc_split_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int HW, int planes) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* s = src + (size_t)b * C * HW;
    __nv_bfloat16* d = dst + (size_t)b * HW * C;
    const size_t plane = (size_t)gridDim.z * HW * C;
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, p = p0 + tx;
        tile[i][tx] = (c < C && p < HW) ? s[(size_t)c * HW + p] : 0.0f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int p = p0 + i, c = c0 + tx;
        if (p < HW && c < C) {
            __nv_bfloat16 pl[3];
            split_planes(tile[tx][i], planes, pl);
            for (int q = 0; q < planes; ++q) d[q * plane + (size_t)p * C + c] = pl[q];
        }
    }
}
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows, int64_t cols, int planes) {
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        __nv_bfloat16 pl[3];
        split_planes(src[i], planes, pl);
        for (int q = 0; q < planes; ++q) dst[(r * planes + q) * cols + c] = pl[q];
    }
}

// TemporalPooling 'direct' (layers.py:202-205) over axis 1 of (B,K,inner) bf16.
__global__ void __launch_bounds__(256)
temporal_pool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int K, int64_t inner, int mode) {
    const int b = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= inner) return;
    const __nv_bfloat16* p = x + (size_t)b * K * inner + i;
    float acc = __bfloat162float(p[0]);
    for (int k = 1; k < K; ++k) {
        float v = __bfloat162float(p[(size_t)k * inner]);
        acc = (mode == VD_JOIN_MAX) ? fmaxf(acc, v) : __fadd_rn(acc, v);
    }
    if (mode == VD_JOIN_MEAN) acc = __fdiv_rn(acc, (float)K);
    y[(size_t)b * inner + i] = __float2bfloat16_rn(acc);
}

}  // namespace vd

using namespace vd;

extern "C" int vd_yolo_decode(const float* pred, int B, int H, int W, int num_class, int num_anchors,
                              const float* anchors_host, float stride, int mode,
                              float* det, int64_t det_rows_total, int64_t det_row_offset,
                              float* raw_centers, float* raw_scales, float* objness, float* class_pred,
                              void* stream) {
    VD_CHECK_ARG(anchors_host && (B == 0 || (pred && det)), "yolo_decode: null pointer");
    VD_CHECK_ARG(B >= 0 && H > 0 && W > 0 && num_class > 0, "yolo_decode: bad shape");
    VD_CHECK_ARG(H <= 128 && W <= 128, "yolo_decode: feature map %dx%d exceeds alloc_size (128,128) (yolo3.py:44)", H, W);
    VD_CHECK_ARG(num_anchors >= 1 && num_anchors <= 6, "yolo_decode: num_anchors %d not in 1..6", num_anchors);
    VD_CHECK_ARG(mode >= 0 && mode <= 2, "yolo_decode: bad mode %d", mode);
    VD_CHECK_ARG(B <= 65535, "yolo_decode: batch > 65535");
    if (mode == VD_MODE_TRAIN && B > 0) VD_CHECK_ARG(raw_centers && raw_scales && objness && class_pred, "yolo_decode: train mode needs the four raw outputs");
    if (B == 0) return VD_OK;
    DecodeArgs a;
    a.pred = pred; a.B = B; a.H = H; a.W = W; a.C = num_class; a.A = num_anchors; a.stride = stride; a.mode = mode;
    for (int i = 0; i < 2 * num_anchors; ++i) a.anchors[i] = anchors_host[i];
    a.det = det; a.det_rows_total = det_rows_total; a.det_row_offset = det_row_offset;
    a.raw_centers = raw_centers; a.raw_scales = raw_scales; a.objness = objness; a.class_pred = class_pred;
    dim3 grid(ceil_div(H * W, 256), B);
    yolo_decode_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

extern "C" int vd_repack_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W, void* stream) {
    VD_CHECK_ARG((B == 0 || (src && dst)) && B >= 0 && C > 0 && H > 0 && W > 0, "repack: bad argument");
    VD_CHECK_ARG(B <= 65535, "repack: batch > 65535");
    if (B == 0) return VD_OK;
    int HW = H * W;
    dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B);
    repack_nchw_to_nhwc_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, C, HW);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

extern "C" int vd_repack_nchw_f32_to_nhwc_split(const float* src, void* dst, int B, int C, int H, int W, int planes, void* stream) {
    VD_CHECK_ARG((B == 0 || (src && dst)) && B >= 0 && C > 0 && H > 0 && W > 0 && (planes == 2 || planes == 3), "repack_split: bad argument");
    VD_CHECK_ARG(B <= 65535, "repack_split: batch > 65535");
    if (B == 0) return VD_OK;
    int HW = H * W;
    dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B);
    repack_nchw_to_nhwc_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, C, HW, planes);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

extern "C" int vd_split_f32_rows(const float* src, void* dst, int64_t rows, int64_t cols, int planes, void* stream) {
    VD_CHECK_ARG(rows >= 0 && cols > 0 && (rows == 0 || (src && dst)) && (planes == 2 || planes == 3), "split_rows: bad argument");
    if (rows == 0) return VD_OK;
    int64_t blocks = ceil_div64(rows * cols, 256); const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    split_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, rows, cols, planes);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

extern "C" int vd_temporal_pool(const void* x, void* y, int B, int K, int64_t inner, int mode, void* stream) {
    VD_CHECK_ARG((B == 0 || (x && y)) && B >= 0 && K >= 1 && inner > 0, "temporal_pool: bad argument");
    VD_CHECK_ARG(mode == VD_JOIN_MAX || mode == VD_JOIN_MEAN, "temporal_pool: mode must be VD_JOIN_MAX or VD_JOIN_MEAN");
    VD_CHECK_ARG(B <= 65535, "temporal_pool: batch > 65535");
    if (B == 0) return VD_OK;
    dim3 grid((unsigned)ceil_div64(inner, 256), B);
    temporal_pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, K, inner, mode);
    VD_LAUNCH_CHECK();
    return VD_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// vd_postprocess_detections -- the host loop of detect() (detect_yolo3.py:222-261) on device: clip the boxes to
// [0, S] (:226), keep the rows with id >= 0 in order (:256), normalise by S (:257), truncate the id (:258), pack
// [id, score, x1, y1, x2, y2] and count the rows per image.  One warp per image (ballot compaction keeps the order).
// ---------------------------------------------------------------------------------------------------------------
namespace vd {
__global__ void __launch_bounds__(256)
postprocess_kernel(const float* __restrict__ ids, const float* __restrict__ scores, const float* __restrict__ bboxes,
                   int frames, int post, float size, float* __restrict__ rows, int32_t* __restrict__ counts) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= frames) return;
    int n = 0;
    for (int j0 = 0; j0 < post; j0 += 32) {
        const int j = j0 + lane;
        const float id = j < post ? ids[(size_t)f * post + j] : -1.0f;
        const bool keep = id >= 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            const float4 b = reinterpret_cast<const float4*>(bboxes)[(size_t)f * post + j];
            float* o = rows + ((size_t)f * post + pos) * 6;
            auto cl = [&](float x) { return x > size ? size : (x < 0.0f ? 0.0f : x); };     // mshadow_op::clip
            o[0] = truncf(id); o[1] = scores[(size_t)f * post + j];
            o[2] = __fdiv_rn(cl(b.x), size); o[3] = __fdiv_rn(cl(b.y), size);
            o[4] = __fdiv_rn(cl(b.z), size); o[5] = __fdiv_rn(cl(b.w), size);
        }
        n += __popc(m);
    }
    for (int j = n * 6 + lane; j < post * 6; j += 32) rows[(size_t)f * post * 6 + j] = -1.0f;
    if (lane == 0) counts[f] = n;
}
}  // namespace vd

extern "C" int vd_postprocess_detections(const float* ids, const float* scores, const float* bboxes, int frames, int post,
                                         float size, float* rows, int32_t* counts, void* stream_) {
    VD_CHECK_ARG(frames >= 0 && post > 0, "postprocess: bad shape frames=%d post=%d", frames, post);
    VD_CHECK_ARG(frames == 0 || (ids && scores && bboxes && rows && counts), "postprocess: null pointer");
    VD_CHECK_ARG(((uintptr_t)bboxes & 15) == 0, "postprocess: bboxes must be 16-byte aligned");
    VD_CHECK_ARG(size > 0.0f, "postprocess: image size must be positive");
    if (frames == 0) return VD_OK;
    vd::postprocess_kernel<<<vd::ceil_div(frames, 8), 256, 0, (cudaStream_t)stream_>>>(ids, scores, bboxes, frames, post, size, rows, counts);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// vd_hierarchical_nms -- detect_yolo3.py:736-789 (`hierarchical_nms`, applied to the predictions of the combined
// class tree at :898-899) on device.  Per image: boxes sorted by class index descending (stable, :756), each lifted to
// its ancestor at `level_thresh` (:765-766), then greedily merged into the kept list: the kept box with the largest
// IoU above `ov_thresh` (first one on ties, :771-775) decides -- none: append; not on the same branch: append; same
// class: max the confidences; otherwise (a descendant already stands there) drop (:777-787).  `iou` is the reference's
// PASCAL-style one (`+ 1` extents, :712-733).  The walk over the boxes is inherently sequential; one warp per image
// runs it, the lanes share the search over the kept list.  arith 0: float64 on the fp32 inputs (= the reference on
// predictions re-loaded from its .txt files, Python floats; pinned by tests/golden/hier_nms_golden.npz); arith 1: the
// legacy-NumPy in-memory path (np.float32 scalars: the six coordinate differences round to fp32, everything after the
// first `+ 1` is float64).
// ---------------------------------------------------------------------------------------------------------------
namespace vd {
constexpr int kHierMaxPost = 256;
constexpr int kHierWarps = 4;

template <int ARITH>
__device__ __forceinline__ double hier_iou(const float4 a, const float4 b) {
    const float x2 = fminf(a.z, b.z), x1 = fmaxf(a.x, b.x), y2 = fminf(a.w, b.w), y1 = fmaxf(a.y, b.y);
    auto diff = [](float hi, float lo) -> double {
        if (ARITH == 1) return (double)__fsub_rn(hi, lo);
        return __dsub_rn((double)hi, (double)lo);
    };
    const double iw = __dadd_rn(diff(x2, x1), 1.0), ih = __dadd_rn(diff(y2, y1), 1.0);
    if (!(iw > 0.0 && ih > 0.0)) return 0.0;
    const double inter = __dmul_rn(iw, ih);
    const double aa = __dmul_rn(__dadd_rn(diff(a.z, a.x), 1.0), __dadd_rn(diff(a.w, a.y), 1.0));
    const double ab = __dmul_rn(__dadd_rn(diff(b.z, b.x), 1.0), __dadd_rn(diff(b.w, b.y), 1.0));
    const double ua = __dsub_rn(__dadd_rn(aa, ab), inter);
    return __ddiv_rn(inter, ua);
}

template <int ARITH>
__global__ void __launch_bounds__(kHierWarps * 32)
hier_nms_kernel(const float* __restrict__ rows, const int32_t* __restrict__ counts, int frames, int post, int C,
                const int32_t* __restrict__ levels, const int32_t* __restrict__ parent, const uint8_t* __restrict__ branch,
                double ov_thresh, double conf_thresh, int level_thresh, float* __restrict__ out_rows, int32_t* __restrict__ out_counts) {
    __shared__ uint16_t s_order[kHierWarps][kHierMaxPost];
    __shared__ int s_kcls[kHierWarps][kHierMaxPost];
    __shared__ float s_kconf[kHierWarps][kHierMaxPost];
    __shared__ float4 s_kbox[kHierWarps][kHierMaxPost];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kHierWarps + w;
    if (f >= frames) return;
    const float* R = rows + (size_t)f * post * 6;
    int n = counts[f]; n = n < 0 ? 0 : (n > post ? post : n);
    // stable descending sort by class index: rank = #{class greater} + #{equal class, earlier row}
    for (int i = lane; i < n; i += 32) {
        const int ci = (int)R[i * 6];
        int rank = 0;
        for (int j = 0; j < n; ++j) { const int cj = (int)R[j * 6]; rank += (cj > ci || (cj == ci && j < i)) ? 1 : 0; }
        s_order[w][rank] = (uint16_t)i;
    }
    __syncwarp();
    int m = 0;
    for (int r = 0; r < n; ++r) {
        const int i = s_order[w][r];
        int cls = (int)R[i * 6];
        const float conf = R[i * 6 + 1];
        if ((double)conf < conf_thresh) continue;
        if (cls < 0 || cls >= C) continue;                                   // not a class of the tree (the reference would raise)
        for (int hop = 0; hop < C && levels[cls] > level_thresh && parent[cls] >= 0; ++hop) cls = parent[cls];   // bounded: inconsistent tables must not hang the GPU
        const float4 bx = make_float4(R[i * 6 + 2], R[i * 6 + 3], R[i * 6 + 4], R[i * 6 + 5]);
        double best = 0.0; int bidx = -1;
        for (int k = lane; k < m; k += 32) {
            const double ov = hier_iou<ARITH>(bx, s_kbox[w][k]);
            if (ov > ov_thresh && ov > best) { best = ov; bidx = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (oi >= 0 && (bidx < 0 || ob > best || (ob == best && oi < bidx))) { best = ob; bidx = oi; }
        }
        bool append = bidx < 0;
        if (!append) {
            const int kc = s_kcls[w][bidx];
            if (!branch[(size_t)cls * C + kc]) append = true;
            else if (cls == kc && lane == 0) s_kconf[w][bidx] = fmaxf(s_kconf[w][bidx], conf);
        }
        if (append) {
            if (lane == 0) { s_kcls[w][m] = cls; s_kconf[w][m] = conf; s_kbox[w][m] = bx; }
            ++m;
        }
        __syncwarp();
    }
    float* O = out_rows + (size_t)f * post * 6;
    for (int k = lane; k < post; k += 32) {
        float v[6] = {-1.f, -1.f, -1.f, -1.f, -1.f, -1.f};
        if (k < m) { const float4 b = s_kbox[w][k]; v[0] = (float)s_kcls[w][k]; v[1] = s_kconf[w][k]; v[2] = b.x; v[3] = b.y; v[4] = b.z; v[5] = b.w; }
#pragma unroll
        for (int e = 0; e < 6; ++e) O[k * 6 + e] = v[e];
    }
    if (lane == 0) out_counts[f] = m;
}
}  // namespace vd

extern "C" int vd_hierarchical_nms(const float* rows, const int32_t* counts, int frames, int post, int num_class,
                                   const int32_t* levels, const int32_t* parent, const uint8_t* branch,
                                   double ov_thresh, double conf_thresh, int level_thresh, int arith,
                                   float* out_rows, int32_t* out_counts, void* stream_) {
    VD_CHECK_ARG(frames >= 0 && post > 0 && num_class > 0, "hierarchical_nms: bad shape frames=%d post=%d classes=%d", frames, post, num_class);
    if (post > vd::kHierMaxPost) return vd::set_error(VD_ERR_UNSUPPORTED, "hierarchical_nms: post = %d rows per image, at most %d", post, vd::kHierMaxPost);
    VD_CHECK_ARG(levels && parent && branch && (frames == 0 || (rows && counts && out_rows && out_counts)), "hierarchical_nms: null pointer");
    VD_CHECK_ARG(arith == 0 || arith == 1, "hierarchical_nms: arith must be 0 (float64) or 1 (legacy float32 scalars)");
    VD_CHECK_ARG(frames == 0 || rows != out_rows, "hierarchical_nms: in-place operation is not supported");
    if (frames == 0) return VD_OK;
    if (level_thresh < 0) level_thresh = 0;                                  // detect_yolo3.py:749
    const int grid = vd::ceil_div(frames, vd::kHierWarps);
    if (arith == 0)
        vd::hier_nms_kernel<0><<<grid, vd::kHierWarps * 32, 0, (cudaStream_t)stream_>>>(rows, counts, frames, post, num_class, levels, parent, branch,
                                                                                    ov_thresh, conf_thresh, level_thresh, out_rows, out_counts);
    else
        vd::hier_nms_kernel<1><<<grid, vd::kHierWarps * 32, 0, (cudaStream_t)stream_>>>(rows, counts, frames, post, num_class, levels, parent, branch,
                                                                                    ov_thresh, conf_thresh, level_thresh, out_rows, out_counts);
    VD_LAUNCH_CHECK();
    return VD_OK;
}
