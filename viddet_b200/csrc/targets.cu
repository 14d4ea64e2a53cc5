// vd_prefetch_targets -- YOLOV3PrefetchTargetGenerator.forward (yolo_target.py:31-148) on device.
//
// The reference allocates (B, sum HW, 9, .) tensors, fills them from a Python double loop with
// one NDArray scalar write per element (yolo_target.py:104-130) and then slices/concats
// (:139-148).  Here:
//   targets_fill_kernel     streams the background (0 / -1) of all five outputs, already in the
//                           final `_slice`d layout, with 128-bit streaming stores  [HBM-write bound:
//                           N*(7+C)*4 bytes per image]
//   targets_scatter_kernel  one CTA per image: fp32 anchor-IoU argmax (box_iou + argmax semantics
//                           of SURVEY.md A.3), fp64 cell index (A.4), `break` at the first invalid
//                           GT, last-writer-wins de-duplication, then the <= M owner rows are
//                           written (class row replaced, not OR-ed).
#include <stdlib.h>
#include "common.cuh"

namespace vd {

struct FillSeg { float* p; size_t n; float v; };
struct FillArgs { FillSeg seg[5]; };

// Background fill through the TMA bulk-copy engine: every CTA keeps one 32 KB shared-memory tile of its segment's value and
// streams it to consecutive 32 KB chunks of the output with cp.async.bulk (shared -> global), one elected thread issuing, up to
// kFillInFlight copies in flight -- a handful of instructions per 32 KB instead of 2048 128-bit stores.  The source tile is never
// modified, so groups are only drained before the CTA exits.
constexpr int kFillTileBytes = 32 * 1024;
constexpr int kFillInFlight = 8;
__global__ void __launch_bounds__(256)
targets_fill_bulk_kernel(FillArgs a) {
    __shared__ __align__(128) float tile[kFillTileBytes / 4];
    const FillSeg s = a.seg[blockIdx.y];
    for (int i = threadIdx.x; i < kFillTileBytes / 4; i += blockDim.x) tile[i] = s.v;
    // unaligned head / tail elements (outputs of torch allocations are 256-byte aligned: normally none)
    size_t head = ((16 - ((uintptr_t)s.p & 15)) & 15) / 4; if (head > s.n) head = s.n;
    const size_t body = (s.n - head) / 4 * 4;                      // floats in whole 16-byte units
    if (blockIdx.x == 0) {
        for (size_t j = threadIdx.x; j < head; j += blockDim.x) s.p[j] = s.v;
        for (size_t j = head + body + threadIdx.x; j < s.n; j += blockDim.x) s.p[j] = s.v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes of the tile -> visible to the bulk-copy engine
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t bytes = body * 4;
        const size_t n_chunks = (bytes + kFillTileBytes - 1) / kFillTileBytes;
        unsigned char* base = reinterpret_cast<unsigned char*>(s.p + head);
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            const size_t off = c * kFillTileBytes;
            const uint32_t nb = (uint32_t)((bytes - off) < (size_t)kFillTileBytes ? (bytes - off) : (size_t)kFillTileBytes);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(src), "r"(nb) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kFillInFlight - 1) : "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(256)
targets_fill_kernel(FillArgs a) {
    const FillSeg s = a.seg[blockIdx.y];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t head = ((16 - ((uintptr_t)s.p & 15)) & 15) / 4; if (head > s.n) head = s.n;
    if (i0 < head) s.p[i0] = s.v;
    float4* p4 = reinterpret_cast<float4*>(s.p + head);
    const size_t n4 = (s.n - head) / 4;
    const float4 v4 = make_float4(s.v, s.v, s.v, s.v);
    size_t i = i0;
    for (; i + 3 * stride < n4; i += 4 * stride) {          // 4 independent 128-bit stores in flight
        __stcs(p4 + i, v4); __stcs(p4 + i + stride, v4); __stcs(p4 + i + 2 * stride, v4); __stcs(p4 + i + 3 * stride, v4);
    }
    for (; i < n4; i += stride) __stcs(p4 + i, v4);
    for (size_t j = head + n4 * 4 + i0; j < s.n; j += stride) s.p[j] = s.v;
}

struct TargetArgs {
    int B, M, C, orig_h, orig_w, ids_width;
    int H[3], W[3], cell_off[4];      // cell_off = cumulative HW (the reference's `_offsets`)
    float aw[9], ah[9];
    const float* gt_boxes; const float* gt_ids; const float* mix;
    float* obj; float* ctr; float* scl; float* wgt; float* cls;
    int32_t* match_out; int32_t* row_out;
    int N;                             // 3 * sum HW
};

constexpr int kTgtThreads = 256;
constexpr int kTgtMaxM = 1024;

__global__ void __launch_bounds__(kTgtThreads)
targets_scatter_kernel(TargetArgs a) {
    __shared__ int s_row[kTgtMaxM];
    __shared__ unsigned char s_match[kTgtMaxM];
    __shared__ unsigned char s_own[kTgtMaxM];
    __shared__ unsigned short s_list[kTgtMaxM];
    __shared__ int s_nv, s_cnt;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) s_nv = a.M;
    __syncthreads();
    const float* gb = a.gt_boxes + (size_t)b * a.M * 4;
    // valid_gts = prod(gt_boxes >= 0) (yolo_target.py:95); the loop BREAKs at the first invalid (:106)
    for (int m = tid; m < a.M; m += kTgtThreads) {
        float4 g = *reinterpret_cast<const float4*>(gb + 4 * m);
        bool valid = (g.x >= 0.f) && (g.y >= 0.f) && (g.z >= 0.f) && (g.w >= 0.f);
        if (!valid) atomicMin(&s_nv, m);
    }
    __syncthreads();
    const int nv = s_nv;
    for (int m = tid; m < a.M; m += kTgtThreads) {
        int row = -1, match = -1;
        if (m < nv) {
            float4 g = *reinterpret_cast<const float4*>(gb + 4 * m);
            // BBoxCornerToCenter (fp32)
            float gw = __fsub_rn(g.z, g.x), gh = __fsub_rn(g.w, g.y);
            float gx = __fadd_rn(g.x, __fdiv_rn(gw, 2.0f)), gy = __fadd_rn(g.y, __fdiv_rn(gh, 2.0f));
            // shifted GT box (-0.5w,-0.5h,0.5w,0.5h) vs zero-centred anchors; box_iou + first argmax
            float r0 = __fmul_rn(-0.5f, gw), r1 = __fmul_rn(-0.5f, gh), r2 = __fmul_rn(0.5f, gw), r3 = __fmul_rn(0.5f, gh);
            float area_r = __fmul_rn(__fsub_rn(r2, r0), __fsub_rn(r3, r1));
            float best = -1.0f;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                float hx = __fdiv_rn(a.aw[j], 2.0f), hy = __fdiv_rn(a.ah[j], 2.0f);
                float l0 = __fsub_rn(0.0f, hx), l1 = __fsub_rn(0.0f, hy), l2 = __fadd_rn(0.0f, hx), l3 = __fadd_rn(0.0f, hy);
                float area_l = __fmul_rn(__fsub_rn(l2, l0), __fsub_rn(l3, l1));
                float iw = __fsub_rn(fminf(l2, r2), fmaxf(l0, r0)); iw = iw > 0.f ? iw : 0.f;
                float ih = __fsub_rn(fminf(l3, r3), fmaxf(l1, r1)); ih = ih > 0.f ? ih : 0.f;
                float inter = __fmul_rn(iw, ih);
                float iou = (inter <= 0.f) ? 0.f : __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_l, area_r), inter));
                if (iou > best) { best = iou; match = j; }          // first maximum wins
            }
            if (match < 0) match = 0;                               // all-NaN row: argmax returns 0
            const int layer = match / 3;
            const int H = a.H[layer], W = a.W[layer];
            // legacy NumPy: np.float32 / python int -> float64
            double fx = (double)gx / (double)a.orig_w * (double)W;
            double fy = (double)gy / (double)a.orig_h * (double)H;
            int loc_x = (int)fx, loc_y = (int)fy;                   // trunc toward 0
            int cell = loc_y * W + loc_x;
            row = 3 * a.cell_off[layer] + cell * 3 + (match - 3 * layer);
            // A cell index past the matched layer's map (centre on / beyond the bottom border: loc_y >= H) lands in another
            // layer's rows of the reference's (B, sum HW, 9, .) scratch, in columns `_slice` discards (yolo_target.py:139-148):
            // nothing visible is written.  (loc_x >= W with the cell still inside the layer IS visible there and is kept.)
            if (cell < 0 || cell >= H * W || row < 0 || row >= a.N) row = -1;
            s_row[m] = row; s_match[m] = (unsigned char)match;
        } else { s_row[m] = -1; s_match[m] = 0; }
        s_own[m] = 0;
        if (a.match_out) a.match_out[(size_t)b * a.M + m] = (m < nv && row >= 0) ? match : -1;
        if (a.row_out) a.row_out[(size_t)b * a.M + m] = row;
    }
    __syncthreads();
    // owners: GT m owns its row iff no later valid GT maps to the same row (last writer wins)
    for (int m = tid; m < nv; m += kTgtThreads) {
        int row = s_row[m];
        if (row < 0) continue;
        bool owner = true;
        for (int m2 = m + 1; m2 < nv; ++m2) if (s_row[m2] == row) { owner = false; break; }
        if (!owner) continue;
        s_own[m] = 1;
        float4 g = *reinterpret_cast<const float4*>(gb + 4 * m);
        float gw = __fsub_rn(g.z, g.x), gh = __fsub_rn(g.w, g.y);
        float gx = __fadd_rn(g.x, __fdiv_rn(gw, 2.0f)), gy = __fadd_rn(g.y, __fdiv_rn(gh, 2.0f));
        int match = s_match[m], layer = match / 3;
        double fx = (double)gx / (double)a.orig_w * (double)a.W[layer];
        double fy = (double)gy / (double)a.orig_h * (double)a.H[layer];
        size_t o = (size_t)b * a.N + row;
        VD_DEV_CHECK(row >= 0 && row < a.N && match >= 0 && match < 9);
        a.ctr[o * 2] = (float)(fx - (double)(int)fx);
        a.ctr[o * 2 + 1] = (float)(fy - (double)(int)fy);
        // np.log(max(gtw, 1) / anchor): fp32 path when gtw >= 1, float64 path when 1 > gtw
        a.scl[o * 2] = !(1.0f > gw) ? logf(__fdiv_rn(gw, a.aw[match])) : (float)log(1.0 / (double)a.aw[match]);
        a.scl[o * 2 + 1] = !(1.0f > gh) ? logf(__fdiv_rn(gh, a.ah[match])) : (float)log(1.0 / (double)a.ah[match]);
        float wv = (float)(2.0 - (double)__fmul_rn(gw, gh) / (double)a.orig_w / (double)a.orig_h);
        a.wgt[o * 2] = wv; a.wgt[o * 2 + 1] = wv;
        a.obj[o] = a.mix ? a.mix[(size_t)b * a.M + m] : 1.0f;
    }
    __syncthreads();
    // class rows of the owners.  (r1: the whole CTA walked the owners one after another, a dependent global load per row: 100 rows
    // x ~1 us = a third of the step at C = 285.)  Now the owners are compacted and the (owner, 128-class chunk) items are spread
    // over the warps, two items (8 independent coalesced loads per lane) in flight per warp.
    if (tid == 0) {
        int n = 0;
        for (int m = 0; m < nv; ++m) if (s_row[m] >= 0 && s_own[m]) s_list[n++] = (unsigned short)m;
        s_cnt = n;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nw = kTgtThreads >> 5;
    const int nchunk = (a.C + 127) >> 7, items = s_cnt * nchunk;
    for (int it0 = warp * 2; it0 < items; it0 += nw * 2) {
        float v[2][4]; float* dst[2]; int c0[2]; bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int item = it0 + u;
            on[u] = item < items;
            const int i = on[u] ? item / nchunk : 0, j = on[u] ? item - i * nchunk : 0;
            const int m = (int)s_list[i];
            c0[u] = j * 128 + lane;
            dst[u] = a.cls + ((size_t)b * a.N + s_row[m]) * a.C;
            if (a.ids_width == 1) {
                int id = (int)a.gt_ids[(size_t)b * a.M + m];
                if (id < 0) id += a.C;                              // numpy negative index
#pragma unroll
                for (int q = 0; q < 4; ++q) v[u][q] = (c0[u] + 32 * q == id) ? 1.0f : 0.0f;
            } else {
                const float* g = a.gt_ids + ((size_t)b * a.M + m) * a.C;
#pragma unroll
                for (int q = 0; q < 4; ++q) v[u][q] = (on[u] && c0[u] + 32 * q < a.C) ? __ldg(g + c0[u] + 32 * q) : 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (on[u] && c0[u] + 32 * q < a.C) dst[u][c0[u] + 32 * q] = v[u][q];
    }
}

}  // namespace vd

using namespace vd;

extern "C" int vd_prefetch_targets(int B, int M, int C, int orig_h, int orig_w, const int* hw_host,
                                   const float* anchors_host, const float* gt_boxes, const float* gt_ids,
                                   int ids_width, const float* mix_or_null,
                                   float* objectness, float* center, float* scale, float* weight, float* cls,
                                   int32_t* match_or_null, int32_t* row_or_null, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VD_CHECK_ARG(B >= 0 && M >= 0 && C > 0 && orig_h > 0 && orig_w > 0, "prefetch_targets: bad shape");
    VD_CHECK_ARG(hw_host && anchors_host, "prefetch_targets: null hw/anchors");
    VD_CHECK_ARG(B == 0 || (objectness && center && scale && weight && cls), "prefetch_targets: null output");
    VD_CHECK_ARG(ids_width == 1 || ids_width == C, "prefetch_targets: gt_ids last dim must be 1 or num_class (%d), got %d", C, ids_width);
    VD_CHECK_ARG(M <= kTgtMaxM, "prefetch_targets: M %d > %d", M, kTgtMaxM);
    VD_CHECK_ARG(B == 0 || M == 0 || (gt_boxes && gt_ids), "prefetch_targets: null gt tensors");
    VD_CHECK_ARG(((uintptr_t)gt_boxes & 15) == 0, "prefetch_targets: gt_boxes must be 16-byte aligned");
    if (B == 0) return VD_OK;
    TargetArgs a;
    a.B = B; a.M = M; a.C = C; a.orig_h = orig_h; a.orig_w = orig_w; a.ids_width = ids_width;
    a.cell_off[0] = 0;
    for (int i = 0; i < 3; ++i) {
        a.H[i] = hw_host[2 * i]; a.W[i] = hw_host[2 * i + 1];
        VD_CHECK_ARG(a.H[i] > 0 && a.W[i] > 0, "prefetch_targets: bad feature map size");
        a.cell_off[i + 1] = a.cell_off[i] + a.H[i] * a.W[i];
    }
    for (int j = 0; j < 9; ++j) { a.aw[j] = anchors_host[2 * j]; a.ah[j] = anchors_host[2 * j + 1]; }
    a.N = 3 * a.cell_off[3];
    a.gt_boxes = gt_boxes; a.gt_ids = gt_ids; a.mix = mix_or_null;
    a.obj = objectness; a.ctr = center; a.scl = scale; a.wgt = weight; a.cls = cls;
    a.match_out = match_or_null; a.row_out = row_or_null;

    FillArgs f;
    const size_t bn = (size_t)B * a.N;
    f.seg[0] = {cls, bn * C, -1.0f};
    f.seg[1] = {objectness, bn, 0.0f};
    f.seg[2] = {center, bn * 2, 0.0f};
    f.seg[3] = {scale, bn * 2, 0.0f};
    f.seg[4] = {weight, bn * 2, 0.0f};
    static const int fill_mode = []() { const char* e = getenv("VD_TARGETS_FILL"); return e ? atoi(e) : 1; }();   // 1: TMA bulk stores (default), 0: per-thread 128-bit streaming stores
    if (fill_mode == 1) {
        // the class tensor is ~98 % of the bytes: give every SM several CTAs there (6 x 32 KB tiles fit an SM), one wave on the small segments
        targets_fill_bulk_kernel<<<dim3(sm_count() * 4, 5), 256, 0, stream>>>(f);
    } else {
        int blocks = sm_count() * 8;
        targets_fill_kernel<<<dim3(blocks, 5), 256, 0, stream>>>(f);
    }
    VD_LAUNCH_CHECK();
    if (M > 0) {
        targets_scatter_kernel<<<B, kTgtThreads, 0, stream>>>(a);
        VD_LAUNCH_CHECK();
    }
    return VD_OK;
}
