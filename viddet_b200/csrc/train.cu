// Training-side consumers of the prefetched targets:
//   vd_target_merge -- YOLOV3DynamicTargetGeneratorSimple + YOLOV3TargetMerger (yolo_target.py:151-281):
//       BBoxBatchIOU(box_preds, gt_boxes) -> max over the GTs -> "ignore" objectness (-1) above ignore_iou_thresh,
//       merged with the prefetched targets wherever the prefetched objectness is positive, plus the class mask.
//   vd_yolo3_loss   -- gluoncv.loss.YOLOV3Loss forward (call site yolo3.py:515): objectness / centre sigmoid-BCE,
//       scale L1, class sigmoid-BCE, each a per-sample mean over the non-batch axes times the element count.
// Both are HBM-bound streaming kernels: one pass over the (B, N, C) class tensors with warp-per-anchor-row coalescing,
// the per-anchor IoU loop runs out of shared memory.  Reductions are two-stage in a fixed order (deterministic).
#include "common.cuh"

namespace vd {

constexpr int kTrainThreads = 256;
constexpr int kTrainTile = 256;            // anchors per block
constexpr int kTrainMaxM = 1024;

__device__ __forceinline__ float mx_max(float a, float b) { return a > b ? a : b; }   // mshadow_op::maximum
__device__ __forceinline__ float mx_min(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float mx_clip(float x, float lo, float hi) { return x > hi ? hi : (x < lo ? lo : x); }

struct MergeArgs {
    int B, N, M, C;
    const float* box_preds; const float* gt_boxes;
    const float* obj_t; const float* centers_t; const float* scales_t; const float* weights_t; const float* clas_t;   // may be null (no prefetched targets)
    float ignore_iou_thresh; int label_smooth; float smooth_weight;
    float* objectness; float* center; float* scale; float* weights; float* class_targets; float* class_mask;
};

__global__ void __launch_bounds__(kTrainThreads)
target_merge_kernel(MergeArgs a) {
    __shared__ float4 s_gt[kTrainMaxM];
    __shared__ float s_area[kTrainMaxM];
    __shared__ unsigned char s_mask[kTrainTile];
    const int b = blockIdx.y, n0 = blockIdx.x * kTrainTile, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    for (int m = tid; m < a.M; m += kTrainThreads) {
        const float4 g = reinterpret_cast<const float4*>(a.gt_boxes)[(size_t)b * a.M + m];
        s_gt[m] = g;
        s_area[m] = __fmul_rn(__fadd_rn(__fsub_rn(g.z, g.x), 0.0f), __fadd_rn(__fsub_rn(g.w, g.y), 0.0f));
    }
    __syncthreads();
    // ---- per anchor: BBoxBatchIOU against every GT (fp32, GluonCV's operation order), max, merge the narrow targets
    const int n = n0 + tid;
    bool mask = false;
    if (n < a.N) {
        const size_t o = (size_t)b * a.N + n;
        const float4 p = reinterpret_cast<const float4*>(a.box_preds)[o];
        const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(p.z, p.x), 0.0f), __fadd_rn(__fsub_rn(p.w, p.y), 0.0f));
        float iou_max = -INFINITY;
        for (int m = 0; m < a.M; ++m) {
            const float4 g = s_gt[m];
            const float iw = mx_clip(__fadd_rn(__fsub_rn(mx_min(p.z, g.z), mx_max(p.x, g.x)), 0.0f), 0.0f, 6.55040e+04f);
            const float ih = mx_clip(__fadd_rn(__fsub_rn(mx_min(p.w, g.w), mx_max(p.y, g.y)), 0.0f), 0.0f, 6.55040e+04f);
            const float inter = __fmul_rn(iw, ih);
            const float uni = __fsub_rn(__fadd_rn(area_a, s_area[m]), inter);
            const float iou = __fdiv_rn(inter, __fadd_rn(uni, 1e-15f));
            iou_max = mx_max(iou_max, iou);
        }
        const float dyn = (iou_max > a.ignore_iou_thresh) ? -1.0f : -0.0f;     // (ious_max > thresh) * -1
        const float ot = a.obj_t ? a.obj_t[o] : 0.0f;
        mask = ot > 0.0f;
        a.objectness[o] = mask ? ot : dyn;
        float2 c = make_float2(0.f, 0.f), s = c, w = c;
        if (mask) {
            c = reinterpret_cast<const float2*>(a.centers_t)[o];
            s = reinterpret_cast<const float2*>(a.scales_t)[o];
            w = reinterpret_cast<const float2*>(a.weights_t)[o];
        }
        reinterpret_cast<float2*>(a.center)[o] = c;
        reinterpret_cast<float2*>(a.scale)[o] = s;
        reinterpret_cast<float2*>(a.weights)[o] = w;
    }
    s_mask[tid] = mask ? 1 : 0;
    __syncthreads();
    // ---- class rows: one warp per anchor row, lanes along the classes (coalesced 128-byte segments)
    const int n_end = min(kTrainTile, a.N - n0);
    for (int r = warp; r < n_end; r += kTrainThreads / 32) {
        const bool mk = s_mask[r] != 0;
        const size_t base = ((size_t)b * a.N + n0 + r) * a.C;
        for (int c = lane; c < a.C; c += 32) {
            float v = mk ? a.clas_t[base + c] : -1.0f;
            if (a.label_smooth) {
                if (v > 0.5f) v = __fsub_rn(v, a.smooth_weight);
                if (!(v < -0.5f || v > 0.5f)) v = a.smooth_weight;
            }
            __stcs(a.class_targets + base + c, v);
            __stcs(a.class_mask + base + c, (mk && v >= 0.0f) ? 1.0f : 0.0f);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct LossArgs {
    int B, N, C, chunks;
    const float* objness; const float* box_centers; const float* box_scales; const float* cls_preds;
    const float* objness_t; const float* center_t; const float* scale_t; const float* weight_t; const float* class_t; const float* class_mask;
    float* partial;            // [B][chunks][4]
    float* out[4];             // 4 x (B)
};

__device__ __forceinline__ float sigmoid_bce(float x, float z, float w) {
    const float relu = x > 0.0f ? x : 0.0f;
    const float soft = log1pf(expf(-fabsf(x)));
    return __fmul_rn(__fadd_rn(__fsub_rn(relu, __fmul_rn(x, z)), soft), w);
}

__device__ __forceinline__ float block_reduce_sum(float v, float* s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.0f;
    for (int w = 0; w < kTrainThreads / 32; ++w) t += s_red[w];       // fixed order
    return t;
}

__global__ void __launch_bounds__(kTrainThreads)
yolo3_loss_partial_kernel(LossArgs a) {
    __shared__ float s_red[kTrainThreads / 32];
    __shared__ float s_objt[kTrainTile];
    const int b = blockIdx.y, n0 = blockIdx.x * kTrainTile, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = n0 + tid;
    float l_obj = 0.f, l_ctr = 0.f, l_scl = 0.f, l_cls = 0.f;
    float ot = 0.0f;
    if (n < a.N) {
        const size_t o = (size_t)b * a.N + n;
        ot = a.objness_t[o];
        const float hard = ot > 0.0f ? 1.0f : ot;
        const float omask = ot > 0.0f ? ot : (ot >= 0.0f ? 1.0f : 0.0f);
        l_obj = sigmoid_bce(a.objness[o], hard, omask);
        const float2 w = reinterpret_cast<const float2*>(a.weight_t)[o];
        const float2 pc = reinterpret_cast<const float2*>(a.box_centers)[o], tc = reinterpret_cast<const float2*>(a.center_t)[o];
        const float2 ps = reinterpret_cast<const float2*>(a.box_scales)[o], ts = reinterpret_cast<const float2*>(a.scale_t)[o];
        const float w0 = __fmul_rn(w.x, ot), w1 = __fmul_rn(w.y, ot);
        l_ctr = sigmoid_bce(pc.x, tc.x, w0) + sigmoid_bce(pc.y, tc.y, w1);
        l_scl = __fmul_rn(fabsf(__fsub_rn(ts.x, ps.x)), w0) + __fmul_rn(fabsf(__fsub_rn(ts.y, ps.y)), w1);
    }
    s_objt[tid] = ot;
    __syncthreads();
    const int n_end = min(kTrainTile, a.N - n0);
    for (int r = warp; r < n_end; r += kTrainThreads / 32) {
        const float otr = s_objt[r];
        const size_t base = ((size_t)b * a.N + n0 + r) * a.C;
        // 4 x 32 classes per step: the eight streaming loads of a step are issued before anything depends on them (r1 issued two
        // loads, then waited: 3.1 TB/s); zero-weight elements contribute x*0 (0 for finite predictions, NaN otherwise, as in the
        // reference): their transcendental work and the label load are skipped.  Same per-lane summation order as before.
        for (int c0 = lane; c0 < a.C; c0 += 128) {
            float cm[4], x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + 32 * q;
                cm[q] = c < a.C ? __ldcs(a.class_mask + base + c) : 0.0f;
                x[q] = c < a.C ? __ldcs(a.cls_preds + base + c) : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + 32 * q;
                if (c < a.C) {
                    const float w = __fmul_rn(cm[q], otr);
                    if (w != 0.0f) l_cls += sigmoid_bce(x[q], __ldcs(a.class_t + base + c), w);
                    else l_cls += __fmul_rn(__fsub_rn(x[q], x[q]), 0.0f);
                }
            }
        }
    }
    float* out = a.partial + ((size_t)b * a.chunks + blockIdx.x) * 4;
    const float so = block_reduce_sum(l_obj, s_red), sc = block_reduce_sum(l_ctr, s_red);
    const float ss = block_reduce_sum(l_scl, s_red), sk = block_reduce_sum(l_cls, s_red);
    if (tid == 0) { out[0] = so; out[1] = sc; out[2] = ss; out[3] = sk; }
}

__global__ void yolo3_loss_final_kernel(LossArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < a.chunks; ++c)
        for (int k = 0; k < 4; ++k) s[k] += (double)a.partial[((size_t)b * a.chunks + c) * 4 + k];
    // mean over the non-batch axes, then times the normaliser (YOLOV3Loss: denorm, denorm*2, denorm*2, denorm_class)
    const float denorm = (float)a.N, denorm_class = (float)a.N * (float)a.C;
    a.out[0][b] = __fmul_rn((float)(s[0] / (double)a.N), denorm);
    a.out[1][b] = __fmul_rn((float)(s[1] / (2.0 * a.N)), denorm * 2.0f);
    a.out[2][b] = __fmul_rn((float)(s[2] / (2.0 * a.N)), denorm * 2.0f);
    a.out[3][b] = __fmul_rn((float)(s[3] / ((double)a.N * a.C)), denorm_class);
}

}  // namespace vd

using namespace vd;

extern "C" int vd_target_merge(int B, int N, int M, int C, const float* box_preds, const float* gt_boxes,
                               const float* obj_t, const float* centers_t, const float* scales_t, const float* weights_t,
                               const float* clas_t, float ignore_iou_thresh, int label_smooth,
                               float* objectness, float* center, float* scale, float* weights, float* class_targets,
                               float* class_mask, void* stream_) {
    VD_CHECK_ARG(B >= 0 && N > 0 && M >= 1 && C >= 1, "target_merge: bad shape B=%d N=%d M=%d C=%d", B, N, M, C);
    VD_CHECK_ARG(M <= kTrainMaxM, "target_merge: M = %d ground-truth boxes per image > %d", M, kTrainMaxM);
    VD_CHECK_ARG(B == 0 || (box_preds && gt_boxes && objectness && center && scale && weights && class_targets && class_mask), "target_merge: null pointer");
    const bool have = obj_t != nullptr;
    VD_CHECK_ARG(!have || (centers_t && scales_t && weights_t && clas_t), "target_merge: prefetched targets must be given together");
    VD_CHECK_ARG((((uintptr_t)box_preds | (uintptr_t)gt_boxes) & 15) == 0, "target_merge: box tensors must be 16-byte aligned");
    if (B == 0) return VD_OK;
    MergeArgs a;
    a.B = B; a.N = N; a.M = M; a.C = C; a.box_preds = box_preds; a.gt_boxes = gt_boxes;
    a.obj_t = obj_t; a.centers_t = centers_t; a.scales_t = scales_t; a.weights_t = weights_t; a.clas_t = clas_t;
    a.ignore_iou_thresh = ignore_iou_thresh; a.label_smooth = label_smooth;
    const float sw = 1.0f / (float)C;
    a.smooth_weight = label_smooth ? (sw < 1.0f / 40.0f ? sw : 1.0f / 40.0f) : sw;
    a.objectness = objectness; a.center = center; a.scale = scale; a.weights = weights; a.class_targets = class_targets; a.class_mask = class_mask;
    dim3 grid((unsigned)ceil_div(N, kTrainTile), (unsigned)B);
    target_merge_kernel<<<grid, kTrainThreads, 0, (cudaStream_t)stream_>>>(a);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

extern "C" size_t vd_yolo3_loss_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return (size_t)B * (size_t)ceil_div(N, kTrainTile) * 4 * sizeof(float);
}

extern "C" int vd_yolo3_loss(int B, int N, int C, const float* objness, const float* box_centers, const float* box_scales,
                             const float* cls_preds, const float* objness_t, const float* center_t, const float* scale_t,
                             const float* weight_t, const float* class_t, const float* class_mask,
                             float* obj_loss, float* center_loss, float* scale_loss, float* cls_loss,
                             void* workspace, size_t workspace_bytes, void* stream_) {
    VD_CHECK_ARG(B >= 0 && N > 0 && C >= 1, "yolo3_loss: bad shape B=%d N=%d C=%d", B, N, C);
    VD_CHECK_ARG(objness && box_centers && box_scales && cls_preds && objness_t && center_t && scale_t && weight_t && class_t && class_mask,
                 "yolo3_loss: null input");
    VD_CHECK_ARG(obj_loss && center_loss && scale_loss && cls_loss, "yolo3_loss: null output");
    if (B == 0) return VD_OK;
    const size_t need = vd_yolo3_loss_workspace_bytes(B, N);
    if (!workspace || workspace_bytes < need) return set_error(VD_ERR_WORKSPACE, "yolo3_loss: workspace %zu < required %zu", workspace_bytes, need);
    LossArgs a;
    a.B = B; a.N = N; a.C = C; a.chunks = ceil_div(N, kTrainTile);
    a.objness = objness; a.box_centers = box_centers; a.box_scales = box_scales; a.cls_preds = cls_preds;
    a.objness_t = objness_t; a.center_t = center_t; a.scale_t = scale_t; a.weight_t = weight_t; a.class_t = class_t; a.class_mask = class_mask;
    a.partial = (float*)workspace;
    a.out[0] = obj_loss; a.out[1] = center_loss; a.out[2] = scale_loss; a.out[3] = cls_loss;
    dim3 grid((unsigned)a.chunks, (unsigned)B);
    yolo3_loss_partial_kernel<<<grid, kTrainThreads, 0, (cudaStream_t)stream_>>>(a);
    VD_LAUNCH_CHECK();
    yolo3_loss_final_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream_>>>(a);
    VD_LAUNCH_CHECK();
    return VD_OK;
}
