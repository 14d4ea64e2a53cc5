/* viddet_b200 -- C ABI of the B200-native VidDet detection-head hot path.
 *
 * The reference (HaydenFaulkner/VidDet) is pure Python on MXNet/Gluon and has no FFI boundary of
 * its own; its "operator API" for this path is the Gluon block / mx.nd.contrib operator call
 * surface.  Each entry point below replaces one such surface and cites it (paths relative to the
 * reference root).  The Python mirror (viddet_b200/blocks.py) binds these with ctypes; the stub a
 * VidDet maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  Every pointer is a DEVICE pointer unless the
 *     name ends in `_host`.  The caller owns every buffer (inputs, outputs, workspace).
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream); no entry point synchronises with the host (mirrors MXNet's async engine).
 *   - Return value: 0 (VD_OK) or a negative VD_ERR_* code; vd_last_error() returns a
 *     thread-local message.  Nothing throws or aborts across the boundary.
 *   - fp32 tensors are dense row-major in the reference's own layouts unless stated.
 *     "NHWC bf16" = channels-last bfloat16 feature map (B, H, W, C): the carrier layout of the
 *     tensor-core path (TMA needs 16-byte strides, which NCHW 13x13/26x26 maps do not have).
 */
#ifndef VIDDET_B200_H_
#define VIDDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VD_OK 0
#define VD_ERR_INVALID_ARG (-1)
#define VD_ERR_UNSUPPORTED (-2)
#define VD_ERR_WORKSPACE   (-3)
#define VD_ERR_CUDA        (-4)

#define VD_MAX_SCALES 3
#define VD_MAX_TOPK   1024      /* largest nms_topk the device NMS handles */
#define VD_MAX_MIRRORS 7        /* peer copies of the fused head's outputs (8 GPUs of one box) */

/* YOLOOutputV3 decode modes (yolo3.py:179-199) */
#define VD_MODE_INFER    0      /* (B, C*HW*A, 6) rows [id, score, x1, y1, x2, y2]            */
#define VD_MODE_TRAIN    1      /* bbox + raw centers/scales/objness/class_pred               */
#define VD_MODE_AGNOSTIC 2      /* (B, HW*A, 6) rows [0, objness, box]                        */

/* temporal join in front of the prediction conv */
#define VD_JOIN_NONE 0
#define VD_JOIN_CAT  1          /* yolo3.py:1134-1136  (B,K,C,H,W)->(B,K*C,H,W)               */
#define VD_JOIN_MAX  2          /* yolo3.py:1137-1138 / layers.py:202-203                      */
#define VD_JOIN_MEAN 3          /* layers.py:204-205                                          */

int vd_version(void);
const char* vd_last_error(void);
/* sm count / compute capability of `device`; fails unless it is an sm_100 part. */
int vd_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * box_nms -- replaces mx.nd.contrib.box_nms as called at yolo3.py:526-528 (x5 sites) and
 * yolo3_temporal.py:545-547.  data/out: (num_batch, num_elem, width) fp32 (leading dims
 * flattened by the caller); record_or_null: (num_batch, num_elem) int32 = MXNet's hidden second
 * output (original row of each kept element, -1 elsewhere).  in/out_format: 0 corner, 1 center.
 * topk <= 0 means num_elem (MXNet's default -1).  Up to VD_MAX_TOPK candidates per image run in shared memory (the reference's call
 * sites: topk = 400); more take the general path (global-memory sort + workspace-resident NMS, O(n * kept) like the operator itself).
 * ------------------------------------------------------------------------------------------ */
size_t vd_box_nms_workspace_bytes(int64_t num_batch, int64_t num_elem, int width, int topk);
int vd_box_nms(const float* data, int64_t num_batch, int64_t num_elem, int width,
               float overlap_thresh, float valid_thresh, int topk, int coord_start,
               int score_index, int id_index, int background_id, int force_suppress,
               int in_format, int out_format, float* out, int32_t* record_or_null,
               void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * YOLOOutputV3 decode -- replaces YOLOOutputV3.hybrid_forward after the conv
 * (yolo3.py:158-199; twin yolo3_temporal.py:139-179).  pred: (B, A*(5+C), H, W) fp32 NCHW (the
 * prediction conv's output).  anchors: 2*A floats (w,h px) on the HOST.
 * VD_MODE_INFER / VD_MODE_AGNOSTIC: writes this scale's rows into det[(b*det_rows_total +
 *   det_row_offset + row)*6 ...], so the three scales can be written straight into the
 *   concatenated tensor of yolo3.py:523.
 * VD_MODE_TRAIN: det = bbox (B, HW*A, 4) (det_rows_total/offset apply likewise, width 4);
 *   raw_centers (B,HW,A,2), raw_scales (B,HW,A,2), objness (B,HW,A,1), class_pred (B,HW,A,C).
 * ------------------------------------------------------------------------------------------ */
int vd_yolo_decode(const float* pred, int B, int H, int W, int num_class, int num_anchors,
                   const float* anchors_host, float stride, int mode,
                   float* det, int64_t det_rows_total, int64_t det_row_offset,
                   float* raw_centers, float* raw_scales, float* objness, float* class_pred,
                   void* stream);

/* Layout helper for reference-layout callers: (B, C, H, W) fp32 -> (B, H, W, C) bf16. */
int vd_repack_nchw_f32_to_nhwc_bf16(const float* src, void* dst_bf16, int B, int C, int H, int W,
                                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Prediction conv -- replaces nn.Conv2D(A*(5+C), 1x1) (yolo3.py:62,157).  tcgen05 GEMM:
 * x NHWC bf16 (B,H,W,Cin) [with join != NONE: (B,K,H,W,Cin)], weight bf16 (N, Cin_total) row-major
 * (= the Gluon (N,Cin,1,1) weight), bias fp32 (N) or NULL -> pred (B, N, H, W) fp32 NCHW.
 * ------------------------------------------------------------------------------------------ */
int vd_pred_conv(const void* x_nhwc_bf16, int B, int H, int W, int Cin, int K_frames, int join,
                 const void* weight_bf16, const float* bias_or_null, int N,
                 float* pred_nchw, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32-parity modes.  The reference head is fp32 end to end (yolo3.py:62,157).  The tensor cores take bf16 operands, so an
 * fp32 value v is carried as P bf16 planes p0 = bf16(v), p1 = bf16(v - p0), p2 = bf16(v - p0 - p1) (8 mantissa bits each:
 * P = 3 holds all 24 bits of v) and the conv accumulates plane products in fp32:
 *   VD_PREC_FP32_SPLIT (P = 3): a0 w0 + a1 w0 + a0 w1 + a1 w1 + a2 w0 + a0 w2  (dropped terms <= 2^-27 relative) -- decoded
 *       scores / boxes within 1e-5 relative of the fp32 reference (tests/test_gpu_fp32.py), 6 bytes per element, 6x the MMAs;
 *   VD_PREC_BF16X2     (P = 2): a0 w0 + a1 w0 + a0 w1  (residual ~ 2^-18 relative) -- within 1e-4, 4 bytes per element, 3x MMAs.
 * Carriers:  activations (P, B, H, W, Cin) bf16, plane-major (every plane is an ordinary channels-last tensor);
 *            weights     (N_out, P, Cin) bf16 (the planes of one output channel side by side).
 * Not combinable with VD_JOIN_CAT; the temporal tip cell has its own entry (vd_temporal_conv_ex) whose split output feeds
 * the head (VD_ERR_UNSUPPORTED otherwise).
 * ------------------------------------------------------------------------------------------ */
#define VD_PREC_BF16        0   /* bf16 operands, fp32 accumulate (1e-3 relative vs the fp32 reference on bf16-representable inputs) */
#define VD_PREC_FP32_SPLIT  1   /* 3 bf16 planes, 6 products, fp32 accumulate (1e-5 relative)                    */
#define VD_PREC_BF16X2      2   /* 2 bf16 planes, 3 products (1e-4 relative)                                     */
/* (B, C, H, W) fp32 -> (planes, B, H, W, C) bf16 planes, planes = 2 or 3 (the activation carrier of the modes above). */
int vd_repack_nchw_f32_to_nhwc_split(const float* src, void* dst_bf16, int B, int C, int H, int W, int planes, void* stream);
/* (rows, cols) fp32 -> (rows, planes, cols) bf16 (the weight carrier of the modes above). */
int vd_split_f32_rows(const float* src, void* dst_bf16, int64_t rows, int64_t cols, int planes, void* stream);
/* vd_pred_conv with a precision selector: VD_PREC_FP32_SPLIT takes the split carriers described above. */
int vd_pred_conv_ex(const void* x, int B, int H, int W, int Cin, int K_frames, int join, int precision,
                    const void* weight, const float* bias_or_null, int N, float* pred_nchw, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused detector tail -- replaces, for one forward call of YOLOV3 / YOLOV3T / YOLOV3Temporal in
 * inference mode: the three YOLOOutputV3 blocks (yolo3.py:496 -> :132-199), the scale concat
 * (:523), box_nms (:526-528), the post_nms slice (:529-530) and the id/score/bbox split
 * (:531-534).  Optionally first applies the temporal tip cell Conv3D((3,1,1))+BN+LeakyReLU
 * (layers.py:82-89, yolo3_temporal.py:226-227) to each (B,T,...) window.
 * ------------------------------------------------------------------------------------------ */
typedef struct VdHeadScale {
    const void* tip_nhwc_bf16;   /* (frames, H, W, Cin) bf16; frames = B (or B*T) ; join!=NONE: (B,K,H,W,Cin) */
    const void* weight_bf16;     /* (N_out, Cin_total) bf16, N_out = A*(5+C), Cin_total = Cin (*K for CAT)  */
    const float* bias;           /* (N_out) fp32 or NULL                                         */
    int H, W, Cin;
    float stride;
    float anchors[6];            /* A = 3 anchors (w,h) px                                       */
    /* optional temporal tip cell in front of the prediction conv (NULL = absent) */
    const void* tconv_weight_bf16;  /* (3, Cin, Cin) bf16: [tap][cout][cin]                      */
    const float* tconv_scale;       /* (Cin) folded BN scale  gamma/sqrt(var+eps)                */
    const float* tconv_shift;       /* (Cin) folded BN shift  beta - mean*scale                  */
    void* tconv_out_nhwc_bf16;      /* (frames, H, W, Cin) bf16 scratch for the cell's output    */
    int tip_window_stride_frames;   /* temporal cell input: 0 / T = materialised windows (B,T,H,W,Cin); 1 = windows sliding over a
                                       resident clip, tip points at window 0's first frame (see vd_temporal_conv_ex) */
    int reserved;                   /* must be 0                                                 */
} VdHeadScale;

/* VdHeadParams.flags */
#define VD_HEAD_NO_FUSED_TIP 1  /* temporal heads: run the tip cell and the head as separate kernels (vd_temporal_conv -> head kernel) even
                                 * where the fused kernel applies (bf16, num_class <= 30 -- 6..19 and 21..29 padded onto the 20 / 30 shapes --, Cin % 256 == 0);
                                 * results are bit-identical */
#define VD_HEAD_NO_PAIR_KERNEL 2 /* wide heads (num_class 31..80): keep the 1-CTA head kernel instead of the CTA-pair one (same results) */
typedef struct VdHeadParams {
    int num_scales;              /* 3, output order s32, s16, s8 (yolo3.py:416-417)             */
    int num_class;
    int frames;                  /* B, or B*T for TimeDistributed heads (layers.py:241-250)     */
    int T;                       /* window length for the temporal cell (frames % T == 0), else 1 */
    int K_frames, join;          /* late temporal join (VD_JOIN_*), K frames per output frame   */
    float nms_thresh;            /* yolo3.py:525: NMS only if 0 < nms_thresh < 1                */
    float valid_thresh;          /* 0.01 at yolo3.py:527                                        */
    int nms_topk;                /* 400 (detect_yolo3.py:200)                                   */
    int post_nms;                /* 100 (yolo3.py:395)                                          */
    int precision;               /* VD_PREC_BF16 (0), or an fp32-parity mode: tips (P,frames,H,W,Cin), weights (N_out,P,Cin) */
    int flags;                   /* 0, or VD_HEAD_* bits below                                   */
    VdHeadScale scale[VD_MAX_SCALES];
    /* Output mirrors (multi-GPU detection gather): every ids / scores / bboxes element vd_head_forward stores at
     * address a is also stored at a + mirror_delta[i] (bytes, multiples of 16), i < n_mirrors.  The targets are the same slot of
     * the gather buffers of the peer GPUs, mapped into this process (vd_ipc_open); 0 mirrors = local outputs only. */
    int n_mirrors;
    int reserved1;               /* must be 0                                                    */
    long long mirror_delta[VD_MAX_MIRRORS];
} VdHeadParams;

/* How the fused call works (and what the workspace is for).  The head kernel filters every frame with a per-frame-slot
 * score threshold left by the previous call and appends the survivors to the frame's candidate list; the NMS kernel
 * proves per frame that the list holds the exact top-k (no overflow, >= k candidates at or above the threshold) and
 * finishes those frames.  Frames that cannot be proven -- all of them on a first call, a few when the data drifts --
 * are redone in the same call by the exact path (per-tile exact top-k selection), which also leaves new thresholds.
 * Results are therefore exact for any input; only the speed depends on the thresholds.
 *
 * Workspace contract.  vd_head_forward keeps state in its workspace BETWEEN calls: those thresholds, warm-start hints of
 * the exact path, the tile-scheduler counters, per-frame histograms / counters (left zeroed by the NMS kernels, so no
 * memset runs per call) and a marker that says so for the layout of the last call.  Hence:
 *   - pass the same buffer to consecutive calls and do not write to it in between (do not share it with
 *     vd_box_nms or other streams' calls);
 *   - a zero-filled buffer, a buffer last used with other parameters, or arbitrary foreign content are all
 *     detected on the device and handled exactly (every frame takes the exact path) -- slower for
 *     that one call, never wrong. */
/* sizeof of the ABI structs as the library was compiled (which: 0 VdHeadScale, 1 VdHeadParams; 0 for anything else):
 * lets a binding check its own struct mirror. */
size_t vd_sizeof(int which);
size_t vd_head_workspace_bytes(const VdHeadParams* p);
/* ids (frames, post_nms, 1), scores (frames, post_nms, 1), bboxes (frames, post_nms, 4) fp32;
 * keep_rows_or_null (frames, post_nms) int32 = row in the (frames, rows, 6) tensor of each output
 * (the NMS keep-indices), -1 for padding. */
int vd_head_forward(const VdHeadParams* p, float* ids, float* scores, float* bboxes,
                    int32_t* keep_rows_or_null, void* workspace, size_t workspace_bytes,
                    void* stream);
/* The same call restricted to some of its kernels (for per-kernel timing with CUDA events on the
 * launching stream; later stages read what earlier stages left in the workspace). */
#define VD_STAGE_TCONV 1        /* temporal tip cell kernels                                   */
#define VD_STAGE_HEAD  2        /* fused pred-conv + decode + candidate-filter kernel          */
#define VD_STAGE_NMS   4        /* list merge passes + per-frame top-k / NMS kernel            */
#define VD_STAGE_ALL   7
int vd_head_forward_stages(const VdHeadParams* p, float* ids, float* scores, float* bboxes,
                           int32_t* keep_rows_or_null, void* workspace, size_t workspace_bytes,
                           void* stream, int stage_mask);
/* Byte offset, inside the workspace, of eight uint32 statistics: word [4] = number of frames of the LAST completed call
 * whose speculative candidate list could not be proven complete and that were redone by the exact path (0 in the steady
 * state; all frames on the first call); word [5] = running total of such frames and word [6] = running total of completed
 * calls on this workspace (both wrap; take differences).  For tests and monitoring; reading needs a stream synchronisation. */
size_t vd_head_stats_offset(const VdHeadParams* p);
/* Byte offset of the profiling-stamp area inside the workspace (VD_DEBUG_HEAD_STAMPS=1: per-tile clock64 of CTA 0; scripts/spec_stamps.py). */
size_t vd_head_debug_offset(const VdHeadParams* p);
/* Number of kernels one vd_head_forward call launches for these parameters (-1 on bad params). */
int vd_head_launch_count(const VdHeadParams* p);
/* 1 if vd_head_forward runs the temporal tip cell and the head as ONE kernel per scale for these parameters (the tip tile stays in
 * shared memory between the two GEMMs: yolo3_temporal.py:226-227 + yolo3.py:157-199 fused), 0 if as separate kernels, -1 on bad params. */
int vd_head_fused_tip(const VdHeadParams* p);
/* Host-side work plan of that kernel for `pairs` CTA pairs (<= 80), for tests and tooling -- no device work: items_out[s] = items (pairs
 * of 128-row tiles) of scale s; pair c runs strided_out[s] rounds of item = round * pairs + c, then the contiguous items
 * beg_out[s * 81 + c] .. beg_out[s * 81 + c + 1]; together the pairs cover every item exactly once.  items_out / strided_out: 3 ints,
 * beg_out: 3 * 81 ints. */
int vd_head_fused_tip_plan(const VdHeadParams* p, int pairs, int* items_out, int* strided_out, int* beg_out);
/* Same conv + decode, but materialises the reference's (frames, rows, 6) detection tensor
 * (what `concat(all_detections)` holds at yolo3.py:523) instead of running NMS. */
int vd_head_detections(const VdHeadParams* p, float* det, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Peer-mapped buffers for the output mirrors above (one process per GPU on one NVLink / NVSwitch box; detect_yolo3.py:211-213
 * shards the batch over the GPUs, the results meet on the host -- here in every rank's gather buffer).  vd_ipc_alloc:
 * cudaMalloc'ed zero-filled device buffer + its 64-byte CUDA IPC handle (exchange it through torch.distributed); vd_ipc_open
 * maps a peer's buffer (peer access enabled lazily); vd_ipc_close unmaps it, vd_ipc_free releases an own buffer.
 * ------------------------------------------------------------------------------------------ */
int vd_ipc_alloc(size_t bytes, void** dev_ptr_out, unsigned char handle_out[64]);
int vd_ipc_open(const unsigned char handle[64], void** dev_ptr_out);
int vd_ipc_close(void* dev_ptr);
int vd_ipc_free(void* dev_ptr);

/* Temporal tip cell alone: Conv3D((3,1,1), pad (1,0,0), no bias) + BN + LeakyReLU(0.1)
 * (layers.py:82-89).  x, y: (B, T, H, W, C) bf16 channels-last; weight (3, C, C) bf16
 * [tap][cout][cin]; scale/shift fp32 (C) folded inference BatchNorm. */
int vd_temporal_conv(const void* x_bf16, void* y_bf16, int B, int T, int H, int W, int C,
                     const void* weight_bf16, const float* scale, const float* shift,
                     float slope, void* stream);
/* The same cell in an fp32-parity mode (see VD_PREC_* above): x, y (P, B, T, H, W, C) bf16 plane-major, weight (P, 3, C, C) bf16
 * [plane][tap][cout][cin]; the fp32 result of BN + LeakyReLU is written as P planes, ready for vd_head_forward with the same
 * precision.  VD_PREC_BF16 is vd_temporal_conv.
 * window_stride_frames: frames between the starts of consecutive windows inside x (0 or T: windows materialised back to back,
 * x = (B, T, H, W, C)).  1 = windows sliding frame by frame over a RESIDENT clip: x points at the first frame of window 0 inside
 * a (L, H, W, C) clip, window b covers clip frames [b, b + T) -- what datasets/imgnetvid.py:480-506 builds for consecutive
 * centres away from the clip's ends (the clamped windows at the ends repeat frames and must be materialised).  The cell's zero
 * padding stays local to each window.  bf16 mode only. */
int vd_temporal_conv_ex(const void* x_bf16, void* y_bf16, int B, int T, int H, int W, int C,
                        const void* weight_bf16, const float* scale, const float* shift,
                        float slope, int precision, int window_stride_frames, void* stream);

/* Conv-BN-LeakyReLU cell of YOLODetectionBlockV3: replaces `_conv2d` (layers.py:63-70) and `_conv3d`
 * (layers.py:73-79) as stacked at yolo3_temporal.py:198-239 -- Conv(no bias, stride 1, zero 'same' padding, kernel extents
 * kt,kh,kw each 1 or 3) + inference BatchNorm (folded into scale/shift) + LeakyReLU(slope).  x (B,T,H,W,Cin), y (B,T,H,W,Cout)
 * bf16 channels-last (T = 1 for 2-D convs); weight (kt*kh*kw, Cout, Cin) bf16 [tap][cout][cin], tap = (it*kh+iy)*kw+ix
 * (cross-correlation order of the Gluon weight (Cout,Cin,kt,kh,kw)); Cin % 64 == 0, Cout % 128 == 0, Cout <= 1024. */
int vd_conv_bn_lrelu(const void* x_bf16, void* y_bf16, int B, int T, int H, int W, int Cin, int Cout,
                     int kt, int kh, int kw, const void* weight_bf16, const float* scale, const float* shift,
                     float slope, void* stream);
/* The glue between two detection blocks (yolo3.py:515-519): nearest x2 upsample of x (`_upsample`, layers.py:10-20), cropped
 * to the route map (`slice_like`), channel-concatenated in front of it.  x (B,H,W,C1), route (B,H2,W2,C2), out (B,H2,W2,C1+C2)
 * bf16 channels-last; H2 <= 2H, W2 <= 2W; C1, C2 multiples of 8. */
int vd_upsample_concat(const void* x_bf16, const void* route_bf16, void* out_bf16, int B, int H, int W, int C1,
                       int H2, int W2, int C2, void* stream);
/* The BW x BH x BF box of pixels x frames (BW*BH*BF <= 128) one MMA tile of vd_conv_bn_lrelu covers on F frames of an
 * H x W map (for roofline maths; F = B*T for kernels without a temporal extent, T otherwise). */
int vd_conv_tile_box(int H, int W, int F, int* BW, int* BH, int* BF);

/* TemporalPooling 'direct' (layers.py:202-205): (B,K,H,W,C) bf16 -> (B,H,W,C) bf16. */
int vd_temporal_pool(const void* x_bf16, void* y_bf16, int B, int K, int64_t inner, int mode,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * YOLOV3PrefetchTargetGenerator.forward -- replaces yolo_target.py:31-148.
 * hw_host: 3x{H,W} of the feature maps in output order; anchors_host: 9x{w,h} in output order
 * (s32,s16,s8).  gt_boxes (B,M,4) corner px padded with -1; gt_ids (B,M,ids_width) with
 * ids_width 1 (class index) or C (multi-hot); mix_or_null (B,M,1).
 * Outputs, already in the `_slice`d final layout (N = 3*sum HW): objectness (B,N,1),
 * center (B,N,2), scale (B,N,2), weight (B,N,2), class (B,N,C).  match/row_or_null (B,M) int32:
 * matched anchor (0..8) and final row of every GT that was written (-1 otherwise).
 * ------------------------------------------------------------------------------------------ */
int vd_prefetch_targets(int B, int M, int C, int orig_h, int orig_w, const int* hw_host,
                        const float* anchors_host, const float* gt_boxes, const float* gt_ids,
                        int ids_width, const float* mix_or_null,
                        float* objectness, float* center, float* scale, float* weight, float* cls,
                        int32_t* match_or_null, int32_t* row_or_null, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-side consumers of the prefetched targets.
 *
 * vd_target_merge -- YOLOV3DynamicTargetGeneratorSimple.hybrid_forward + YOLOV3TargetMerger.hybrid_forward
 * (yolo_target.py:175-205, :226-281; call site yolo3.py:514).  box_preds (B,N,4) corner boxes of the
 * train-mode decode, gt_boxes (B,M,4) padded with -1, prefetched targets obj_t (B,N,1), centers_t /
 * scales_t / weights_t (B,N,2), clas_t (B,N,C) -- all five may be NULL (dynamic targets only).
 * Outputs: objectness (B,N,1) [1 positive / -1 ignored / 0], center, scale, weights (B,N,2),
 * class_targets (B,N,C), class_mask (B,N,C).
 * ------------------------------------------------------------------------------------------ */
int vd_target_merge(int B, int N, int M, int C, const float* box_preds, const float* gt_boxes,
                    const float* obj_t, const float* centers_t, const float* scales_t,
                    const float* weights_t, const float* clas_t, float ignore_iou_thresh,
                    int label_smooth, float* objectness, float* center, float* scale, float* weights,
                    float* class_targets, float* class_mask, void* stream);

/* vd_yolo3_loss -- gluoncv.loss.YOLOV3Loss forward as called at yolo3.py:515: objness (B,N,1),
 * box_centers / box_scales (B,N,2), cls_preds (B,N,C) raw predictions + the six merged targets
 * -> obj_loss, center_loss, scale_loss, cls_loss, each (B).  Deterministic two-stage reduction. */
size_t vd_yolo3_loss_workspace_bytes(int B, int N);
int vd_yolo3_loss(int B, int N, int C, const float* objness, const float* box_centers,
                  const float* box_scales, const float* cls_preds, const float* objness_t,
                  const float* center_t, const float* scale_t, const float* weight_t,
                  const float* class_t, const float* class_mask, float* obj_loss,
                  float* center_loss, float* scale_loss, float* cls_loss, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Post-processing of detect() on device (detect_yolo3.py:222-261): boxes clipped
 * to [0, size], rows with id >= 0 kept in order, boxes divided by size, ids truncated.
 * ids/scores (frames, post, 1), bboxes (frames, post, 4) -> rows (frames, post, 6)
 * [id, score, x1, y1, x2, y2] padded with -1, counts (frames) int32.
 * ------------------------------------------------------------------------------------------ */
int vd_postprocess_detections(const float* ids, const float* scores, const float* bboxes, int frames,
                              int post, float size, float* rows, int32_t* counts, void* stream);

/* `hierarchical_nms` of detect_yolo3.py:736-789 (with its `iou`, :712-733), applied at :898-899 to the predictions of the
 * combined class tree.  rows (frames, post, 6) fp32 [cls, conf, x1, y1, x2, y2] with counts (frames) valid rows per image
 * (the layout vd_postprocess_detections writes); levels (C) = dataset.get_levels() (combined.py:117-126); parent (C) = class
 * index of the parent, -1 under ROOT (:766); branch (C, C) uint8 = dataset.on_branch(i, j) (combined.py:143-150).  arith 0:
 * float64 arithmetic on the fp32 inputs (the reference on re-loaded predictions); 1: legacy-NumPy float32-scalar path.
 * out_rows (frames, post, 6) padded with -1, out_counts (frames).  post <= 256. */
int vd_hierarchical_nms(const float* rows, const int32_t* counts, int frames, int post, int num_class,
                        const int32_t* levels, const int32_t* parent, const uint8_t* branch,
                        double ov_thresh, double conf_thresh, int level_thresh, int arith,
                        float* out_rows, int32_t* out_counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDDET_B200_H_ */
