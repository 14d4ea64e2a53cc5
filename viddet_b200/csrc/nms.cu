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

// ---------------------------------------------------------------------------------------------------------------
// More than VD_MAX_TOPK candidates per image (MXNet's default topk = -1 on a long input, or a large explicit topk): the
// general path.  Every element gets a 64-bit key (0 = invalid), each image's keys are sorted by a global-memory bitonic
// network (2048-key runs in shared memory, wider strides one pass each), and ONE CTA per image runs the wavefront NMS of the
// fused head (nms_tail_wave) with its per-rank / kept-box arrays in the workspace instead of shared memory.  Like the
// operator it replaces this is O(n * kept) in the worst case: it is the compatibility path, the reference's own call
// sites pass topk = 400 (yolo3.py:526-528).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRun = 2048;                  // keys one CTA sorts in shared memory

__global__ void __launch_bounds__(256)
nms_large_keys_kernel(BoxNmsArgs a, uint64_t* __restrict__ keys, int64_t npad, uint32_t* __restrict__ nvalid) {
    const int b = blockIdx.y;
    const float* in = a.data + (size_t)b * a.num_elem * a.width;
    const bool bg = a.id_index >= 0 && a.background_id >= 0;
    uint32_t cnt = 0;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < npad; row += (int64_t)gridDim.x * blockDim.x) {
        uint64_t key = 0ull;
        if (row < a.num_elem) {
            const float s = __ldg(in + row * a.width + a.score_index);
            bool ok = s > a.valid_thresh;
            if (ok && bg) ok = (int)__ldg(in + row * a.width + a.id_index) != a.background_id;
            if (ok) { key = make_key(s, (uint32_t)row); ++cnt; }
        }
        keys[(size_t)b * npad + row] = key;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&nvalid[b], cnt);
}
// element i of a run that must end up sorted in direction `desc`: keeps the larger of (a, partner) iff it is the lower index (desc)
__device__ __forceinline__ void cmpx(uint64_t& lo, uint64_t& hi, bool desc) {
    if ((lo < hi) == desc) { const uint64_t t = lo; lo = hi; hi = t; }
}
// sorts every run of kRun keys; run r is sorted descending iff bit log2(kRun) of its first index is clear (bitonic convention for the merges that follow)
__global__ void __launch_bounds__(kRun / 2)
nms_large_sort_runs_kernel(uint64_t* __restrict__ keys, int64_t npad) {
    __shared__ uint64_t s[kRun];
    uint64_t* g = keys + (size_t)blockIdx.y * npad + (size_t)blockIdx.x * kRun;
    const int t = threadIdx.x;
    s[t] = g[t]; s[t + kRun / 2] = g[t + kRun / 2];
    const int64_t base = (int64_t)blockIdx.x * kRun;
    for (int size = 2; size <= kRun; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            const int i = 2 * t - (t & (stride - 1)), j = i + stride;
            const bool desc = (((base + i) & size) == 0);
            cmpx(s[i], s[j], desc);
        }
    }
    __syncthreads();
    g[t] = s[t]; g[t + kRun / 2] = s[t + kRun / 2];
}
// one compare-exchange pass of the bitonic merge of width `size` at distance `stride` (>= kRun)
__global__ void __launch_bounds__(256)
nms_large_global_step_kernel(uint64_t* __restrict__ keys, int64_t npad, int64_t size, int64_t stride) {
    uint64_t* g = keys + (size_t)blockIdx.y * npad;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npad / 2; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = 2 * t - (t & (stride - 1)), j = i + stride;
        const bool desc = ((i & size) == 0);
        uint64_t a = g[i], b = g[j];
        if ((a < b) == desc) { g[i] = b; g[j] = a; }
    }
}
// the remaining passes (distances kRun/2 .. 1) of the merge of width `size`, inside shared memory
__global__ void __launch_bounds__(kRun / 2)
nms_large_merge_runs_kernel(uint64_t* __restrict__ keys, int64_t npad, int64_t size) {
    __shared__ uint64_t s[kRun];
    uint64_t* g = keys + (size_t)blockIdx.y * npad + (size_t)blockIdx.x * kRun;
    const int t = threadIdx.x;
    s[t] = g[t]; s[t + kRun / 2] = g[t + kRun / 2];
    const bool desc = ((((int64_t)blockIdx.x * kRun) & size) == 0);
    for (int stride = kRun / 2; stride > 0; stride >>= 1) {
        __syncthreads();
        const int i = 2 * t - (t & (stride - 1)), j = i + stride;
        cmpx(s[i], s[j], desc);
    }
    __syncthreads();
    g[t] = s[t]; g[t + kRun / 2] = s[t + kRun / 2];
}
struct LargeSource {
    CompatSource inner; int class_aware;
    __device__ __forceinline__ void load(int b, uint32_t row, float s, float4& bx, int& c, float& area) const {
        inner.load(b, row, s, bx, c, area);
        if (!class_aware) c = 0;                 // force_suppress / no id column: every pair is comparable
    }
};
struct LargeScratch { float4* sbox; int* scls; float* sarea; float4* skbox; float* skarea; int* skcls; size_t per_image, kept_per_image; };
__global__ void __launch_bounds__(kNmsThreads)
nms_large_kernel(const uint64_t* __restrict__ keys, int64_t npad, const uint32_t* __restrict__ nvalid, NmsParams P, BoxNmsArgs a, LargeScratch w) {
    __shared__ WaveShared wsh;
    const int b = blockIdx.x;
    const uint32_t nv = nvalid[b];
    const int n = (int)(nv < (uint32_t)P.k ? nv : (uint32_t)P.k);
    LargeSource src{CompatSource{a}, P.class_aware};
    CompatSink sink{a};
    nms_tail_wave(n, b, P, src, sink, keys + (size_t)b * npad, w.sbox + (size_t)b * w.per_image, w.scls + (size_t)b * w.per_image,
                  w.sarea + (size_t)b * w.per_image, w.skbox + (size_t)b * w.kept_per_image, w.skarea + (size_t)b * w.kept_per_image,
                  w.skcls + (size_t)b * w.kept_per_image, &wsh);
}
struct LargePlan { int64_t npad; size_t off_keys, off_nvalid, off_sbox, off_scls, off_sarea, off_skbox, off_skarea, off_skcls, per_image, kept_per_image, total; };
static LargePlan nms_large_plan(int64_t num_batch, int64_t num_elem, int64_t k) {
    LargePlan p;
    p.npad = kRun; while (p.npad < num_elem) p.npad <<= 1;
    p.per_image = (size_t)k; p.kept_per_image = (size_t)k + 32 + 4 * (kNmsThreads / 32);
    size_t off = 0;
    p.off_keys = off; off += align_up((size_t)num_batch * p.npad * 8, 256);
    p.off_nvalid = off; off += align_up((size_t)num_batch * 4, 256);
    p.off_sbox = off; off += align_up((size_t)num_batch * p.per_image * 16, 256);
    p.off_scls = off; off += align_up((size_t)num_batch * p.per_image * 4, 256);
    p.off_sarea = off; off += align_up((size_t)num_batch * p.per_image * 4, 256);
    p.off_skbox = off; off += align_up((size_t)num_batch * p.kept_per_image * 16, 256);
    p.off_skarea = off; off += align_up((size_t)num_batch * p.kept_per_image * 4, 256);
    p.off_skcls = off; off += align_up((size_t)num_batch * p.kept_per_image * 4, 256);
    p.total = off;
    return p;
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

extern "C" size_t vd_box_nms_workspace_bytes(int64_t num_batch, int64_t num_elem, int, int topk) {
    if (num_batch <= 0 || num_elem <= 0) return 256;
    const int64_t k64 = (topk > 0 && topk < num_elem) ? topk : num_elem;
    if (k64 > VD_MAX_TOPK) return nms_large_plan(num_batch, num_elem, k64).total;
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
    if (k64 > VD_MAX_TOPK) {
        // general path: global bitonic sort + wavefront NMS with workspace-resident arrays
        VD_CHECK_ARG(k64 < (1ll << 30), "box_nms: too many candidates");
        const LargePlan lp = nms_large_plan(num_batch, num_elem, k64);
        if (!workspace || workspace_bytes < lp.total)
            return set_error(VD_ERR_WORKSPACE, "box_nms: workspace %zu < required %zu", workspace_bytes, lp.total);
        BoxNmsArgs a{data, out, record_or_null, num_elem, width, valid_thresh, coord_start, score_index, id_index, background_id, in_format, out_format};
        unsigned char* ws = (unsigned char*)workspace;
        uint64_t* keys = (uint64_t*)(ws + lp.off_keys);
        uint32_t* nvalid = (uint32_t*)(ws + lp.off_nvalid);
        {
            size_t n_out = (size_t)num_batch * num_elem * width, n_rec = (size_t)num_batch * num_elem;
            int blocks = (int)((n_out / 4 + 255) / 256); int maxb = sm_count() * 16;
            if (blocks > maxb) blocks = maxb; if (blocks < 1) blocks = 1;
            nms_fill_kernel<<<blocks, 256, 0, stream>>>(out, n_out, record_or_null, record_or_null ? n_rec : 0);
            VD_LAUNCH_CHECK();
        }
        VD_CUDA(cudaMemsetAsync(nvalid, 0, (size_t)num_batch * 4, stream));
        int kb = (int)((lp.npad + 255) / 256); if (kb > sm_count() * 8) kb = sm_count() * 8;
        nms_large_keys_kernel<<<dim3((unsigned)kb, (unsigned)num_batch), 256, 0, stream>>>(a, keys, lp.npad, nvalid);
        VD_LAUNCH_CHECK();
        const unsigned runs = (unsigned)(lp.npad / kRun);
        nms_large_sort_runs_kernel<<<dim3(runs, (unsigned)num_batch), kRun / 2, 0, stream>>>(keys, lp.npad);
        VD_LAUNCH_CHECK();
        for (int64_t size = 2 * kRun; size <= lp.npad; size <<= 1) {
            for (int64_t stride = size >> 1; stride >= kRun; stride >>= 1) {
                int gb = (int)((lp.npad / 2 + 255) / 256); if (gb > sm_count() * 8) gb = sm_count() * 8;
                nms_large_global_step_kernel<<<dim3((unsigned)gb, (unsigned)num_batch), 256, 0, stream>>>(keys, lp.npad, size, stride);
                VD_LAUNCH_CHECK();
            }
            nms_large_merge_runs_kernel<<<dim3(runs, (unsigned)num_batch), kRun / 2, 0, stream>>>(keys, lp.npad, size);
            VD_LAUNCH_CHECK();
        }
        NmsParams P;
        P.overlap_thresh = overlap_thresh; P.k = (int)k64; P.sortn = 0;
        P.class_aware = (!force_suppress && id_index >= 0) ? 1 : 0; P.max_out = (int)k64; P.dbg = nullptr;
        LargeScratch w;
        w.sbox = (float4*)(ws + lp.off_sbox); w.scls = (int*)(ws + lp.off_scls); w.sarea = (float*)(ws + lp.off_sarea);
        w.skbox = (float4*)(ws + lp.off_skbox); w.skarea = (float*)(ws + lp.off_skarea); w.skcls = (int*)(ws + lp.off_skcls);
        w.per_image = lp.per_image; w.kept_per_image = lp.kept_per_image;
        nms_large_kernel<<<(unsigned)num_batch, kNmsThreads, 0, stream>>>(keys, lp.npad, nvalid, P, a, w);
        VD_LAUNCH_CHECK();
        return VD_OK;
    }
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
