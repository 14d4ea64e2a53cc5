// vd_box_nms -- drop-in for mx.nd.contrib.box_nms (call sites yolo3.py:526-528 x5,
// yolo3_temporal.py:545-547).  Semantics: SURVEY.md Appendix A.3.
//
// Kernels (all HBM-/latency-bound integer + fp32-compare work, no tensor cores):
//   nms_fill_kernel     out = -1, record = -1                       (streaming 128-bit stores)
//   nms_filter_kernel   one CTA per 16 Ki-row chunk: valid filter, 64-bit keys in registers,
//                       chunk-local top-k pivot, survivors (<= 1024) -> candidate list
//   nms_merge_kernel    (only if > 8 chunks) 8 lists -> 1, repeated
//   nms_final_kernel    one CTA per image: exact top-k, sort, bit-matrix NMS, compaction
#include "nms_core.cuh"

namespace vd {

constexpr int kFilterThreads = 512;
constexpr int kFilterR = 32;
constexpr int kChunkRows = kFilterThreads * kFilterR;       // 16384 rows per CTA

struct BoxNmsArgs {
    const float* data; float* out; int32_t* record;
    int64_t num_elem; int width;
    float valid_thresh; int coord_start, score_index, id_index, background_id;
    int in_format, out_format;
};

__global__ void __launch_bounds__(256)
nms_fill_kernel(float* __restrict__ out, size_t n_out, int32_t* __restrict__ rec, size_t n_rec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // out is cudaMalloc/torch aligned; handle a possibly unaligned head generically
    size_t head = ((16 - ((uintptr_t)out & 15)) & 15) / 4; if (head > n_out) head = n_out;
    if (i0 < head) out[i0] = -1.0f;
    float4* o4 = reinterpret_cast<float4*>(out + head);
    size_t n4 = (n_out - head) / 4;
    const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
    for (size_t i = i0; i < n4; i += stride) __stcs(o4 + i, m1);
    for (size_t i = head + n4 * 4 + i0; i < n_out; i += stride) out[i] = -1.0f;
    if (rec) for (size_t i = i0; i < n_rec; i += stride) rec[i] = -1;
}

__global__ void __launch_bounds__(kFilterThreads, 1)
nms_filter_kernel(BoxNmsArgs a, int k, uint64_t* __restrict__ lists, uint32_t* __restrict__ counts,
                  int n_lists) {
    __shared__ SelectScratch scr;
    __shared__ uint64_t stage[kListCap];
    const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    select_scratch_init(&scr);
    const float* in = a.data + (size_t)b * a.num_elem * a.width;
    const int64_t row0 = (int64_t)chunk * kChunkRows;
    uint64_t keys[kFilterR];
    const bool bg = a.id_index >= 0 && a.background_id >= 0;
#pragma unroll
    for (int r = 0; r < kFilterR; ++r) {
        int64_t row = row0 + (int64_t)r * kFilterThreads + tid;
        uint64_t key = 0ull;
        if (row < a.num_elem) {
            float s = __ldg(in + row * a.width + a.score_index);
            bool ok = s > a.valid_thresh;                                  // strict; NaN invalid
            if (ok && bg) ok = (int)__ldg(in + row * a.width + a.id_index) != a.background_id;
            if (ok) key = make_key(s, (uint32_t)row);
        }
        keys[r] = key;
    }
    int it = 0; uint32_t nsel = 0;
    uint64_t piv = block_select_pivot<kFilterR>(keys, (uint32_t)k, (uint32_t)kListCap, 0ull, 0ull, &scr, it, &nsel);
    block_compact<kFilterR>(keys, piv, stage, (uint32_t)kListCap, &scr.out_count);
    __syncthreads();
    uint64_t* lout = lists + ((size_t)b * n_lists + chunk) * kListCap;
    for (int i = tid; i < (int)nsel; i += blockDim.x) lout[i] = stage[i];
    if (tid == 0) counts[(size_t)b * n_lists + chunk] = nsel;
}

struct CompatSource {
    BoxNmsArgs a;
    __device__ __forceinline__ void load(int b, uint32_t row, float, float4& bx, int& c, float& area) const {
        const float* p = a.data + ((size_t)b * a.num_elem + row) * a.width;
        float v0 = p[a.coord_start], v1 = p[a.coord_start + 1], v2 = p[a.coord_start + 2], v3 = p[a.coord_start + 3];
        if (a.in_format == 0) {
            bx = make_float4(v0, v1, v2, v3);
            area = vd_box_area(__fsub_rn(v2, v0), __fsub_rn(v3, v1));
        } else {                                   // center: MXNet Intersect(): a1 -/+ a2/2
            float hw = __fdiv_rn(v2, 2.0f), hh = __fdiv_rn(v3, 2.0f);
            bx = make_float4(__fsub_rn(v0, hw), __fsub_rn(v1, hh), __fadd_rn(v0, hw), __fadd_rn(v1, hh));
            area = vd_box_area(v2, v3);
        }
        c = a.id_index >= 0 ? (int)p[a.id_index] : 0;
    }
};
struct CompatSink {
    BoxNmsArgs a;
    __device__ __forceinline__ void emit(int b, int pos, uint32_t row, float, float4, int) const {
        const float* p = a.data + ((size_t)b * a.num_elem + row) * a.width;
        float* o = a.out + ((size_t)b * a.num_elem + pos) * a.width;
        for (int c = 0; c < a.width; ++c) o[c] = p[c];
        if (a.in_format != a.out_format) {
            float* q = o + a.coord_start;
            float v0 = p[a.coord_start], v1 = p[a.coord_start + 1], v2 = p[a.coord_start + 2], v3 = p[a.coord_start + 3];
            if (a.out_format == 0) {               // center -> corner
                float hw = __fdiv_rn(v2, 2.0f), hh = __fdiv_rn(v3, 2.0f);
                q[0] = __fsub_rn(v0, hw); q[1] = __fsub_rn(v1, hh); q[2] = __fadd_rn(v0, hw); q[3] = __fadd_rn(v1, hh);
            } else {                               // corner -> center
                q[0] = __fdiv_rn(__fadd_rn(v0, v2), 2.0f); q[1] = __fdiv_rn(__fadd_rn(v1, v3), 2.0f);
                q[2] = __fsub_rn(v2, v0); q[3] = __fsub_rn(v3, v1);
            }
        }
        if (a.record) a.record[(size_t)b * a.num_elem + pos] = (int32_t)row;
    }
    __device__ __forceinline__ void finish(int, int) const {}
};

__global__ void __launch_bounds__(kFinalThreads, 1)
nms_final_compat_kernel(const uint64_t* __restrict__ lists, const uint32_t* __restrict__ counts,
                        int n_lists, NmsParams P, BoxNmsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    CompatSource src{a}; CompatSink sink{a};
    nms_final_body(lists + (size_t)b * n_lists * kListCap, counts + (size_t)b * n_lists, n_lists, b, P,
                   src, sink, smem_raw);
}

// workspace: level-0 lists [NB][n0][cap] + counts, then ping-pong merge levels
struct NmsPlan { int n0; size_t lists_bytes, counts_bytes, total; };
static NmsPlan nms_plan(int64_t num_batch, int64_t num_elem) {
    NmsPlan p;
    p.n0 = (int)ceil_div64(num_elem > 0 ? num_elem : 1, kChunkRows);
    int n1 = ceil_div(p.n0, kMaxLists);
    p.lists_bytes = align_up((size_t)num_batch * p.n0 * kListCap * 8, 256);
    size_t l1 = align_up((size_t)num_batch * n1 * kListCap * 8, 256);
    p.counts_bytes = align_up((size_t)num_batch * p.n0 * 4, 256);
    p.total = p.lists_bytes + 2 * l1 + 3 * p.counts_bytes;
    return p;
}

}  // namespace vd

using namespace vd;

extern "C" size_t vd_box_nms_workspace_bytes(int64_t num_batch, int64_t num_elem, int, int) {
    if (num_batch <= 0 || num_elem <= 0) return 256;
    return nms_plan(num_batch, num_elem).total;
}

extern "C" int vd_box_nms(const float* data, int64_t num_batch, int64_t num_elem, int width,
                          float overlap_thresh, float valid_thresh, int topk, int coord_start,
                          int score_index, int id_index, int background_id, int force_suppress,
                          int in_format, int out_format, float* out, int32_t* record_or_null,
                          void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VD_CHECK_ARG(num_batch >= 0 && num_elem >= 0, "box_nms: negative shape");
    if (num_batch == 0 || num_elem == 0) return VD_OK;
    VD_CHECK_ARG(data && out, "box_nms: null data/out");
    VD_CHECK_ARG(width >= 4 && coord_start >= 0 && coord_start + 4 <= width, "box_nms: coord_start %d + 4 > width %d", coord_start, width);
    VD_CHECK_ARG(score_index >= 0 && score_index < width, "box_nms: score_index %d out of range", score_index);
    VD_CHECK_ARG(id_index < width, "box_nms: id_index %d out of range", id_index);
    VD_CHECK_ARG((in_format == 0 || in_format == 1) && (out_format == 0 || out_format == 1), "box_nms: bad format");
    VD_CHECK_ARG(num_elem < (1ll << 31), "box_nms: num_elem too large");
    VD_CHECK_ARG(num_batch <= 65535, "box_nms: num_batch %lld > 65535", (long long)num_batch);
    int64_t k64 = (topk > 0 && topk < num_elem) ? topk : num_elem;
    if (k64 > VD_MAX_TOPK)
        return set_error(VD_ERR_UNSUPPORTED, "box_nms: min(topk, num_elem) = %lld exceeds VD_MAX_TOPK = %d",
                         (long long)k64, VD_MAX_TOPK);
    const int k = (int)k64;
    NmsPlan plan = nms_plan(num_batch, num_elem);
    if (!workspace || workspace_bytes < plan.total)
        return set_error(VD_ERR_WORKSPACE, "box_nms: workspace %zu < required %zu", workspace_bytes, plan.total);

    BoxNmsArgs a{data, out, record_or_null, num_elem, width, valid_thresh, coord_start, score_index,
                 id_index, background_id, in_format, out_format};
    unsigned char* ws = (unsigned char*)workspace;
    uint64_t* lists0 = (uint64_t*)ws;
    size_t l1 = (plan.total - plan.lists_bytes - 3 * plan.counts_bytes) / 2;
    uint64_t* listsA = (uint64_t*)(ws + plan.lists_bytes);
    uint64_t* listsB = (uint64_t*)(ws + plan.lists_bytes + l1);
    uint32_t* counts0 = (uint32_t*)(ws + plan.lists_bytes + 2 * l1);
    uint32_t* countsA = (uint32_t*)(ws + plan.lists_bytes + 2 * l1 + plan.counts_bytes);
    uint32_t* countsB = (uint32_t*)(ws + plan.lists_bytes + 2 * l1 + 2 * plan.counts_bytes);

    {   // -1 prefill of out / record (A.3 step 6)
        size_t n_out = (size_t)num_batch * num_elem * width, n_rec = (size_t)num_batch * num_elem;
        int blocks = (int)((n_out / 4 + 255) / 256); int maxb = sm_count() * 16;
        if (blocks > maxb) blocks = maxb; if (blocks < 1) blocks = 1;
        nms_fill_kernel<<<blocks, 256, 0, stream>>>(out, n_out, record_or_null, record_or_null ? n_rec : 0);
        VD_LAUNCH_CHECK();
    }
    nms_filter_kernel<<<dim3(plan.n0, (unsigned)num_batch), kFilterThreads, 0, stream>>>(a, k, lists0, counts0, plan.n0);
    VD_LAUNCH_CHECK();
    const uint64_t* lists = lists0; const uint32_t* counts = counts0; int n_lists = plan.n0;
    uint64_t* lout = listsA; uint32_t* cout = countsA;
    while (n_lists > kMaxLists) {
        int n_out = ceil_div(n_lists, kMaxLists);
        nms_merge_kernel<<<dim3(n_out, (unsigned)num_batch), kFinalThreads, 0, stream>>>(lists, counts, n_lists, lout, cout, n_out, k);
        VD_LAUNCH_CHECK();
        lists = lout; counts = cout; n_lists = n_out;
        lout = (lout == listsA) ? listsB : listsA; cout = (cout == countsA) ? countsB : countsA;
    }
    NmsParams P;
    P.overlap_thresh = overlap_thresh; P.k = k; P.sortn = nms_sortn(k);
    P.class_aware = (!force_suppress && id_index >= 0) ? 1 : 0; P.max_out = k; P.dbg = nullptr;
    size_t smem = nms_final_smem(k);
    { int rc_ = configure_kernel((const void*)nms_final_compat_kernel, (int)nms_final_smem(VD_MAX_TOPK), false); if (rc_) return rc_; }
    nms_final_compat_kernel<<<(unsigned)num_batch, kFinalThreads, smem, stream>>>(lists, counts, n_lists, P, a);
    VD_LAUNCH_CHECK();
    return VD_OK;
}
