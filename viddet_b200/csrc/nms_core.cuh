// Per-image exact top-k + class-aware greedy NMS, executed by ONE CTA per image on candidate
// lists produced upstream (box_nms filter kernel or the fused head epilogue).
//
// Replaces the sort + `nms_impl` x topk sequential launches + `nms_assign` inside MXNet's
// _contrib_box_nms (call sites yolo3.py:526-528, yolo3_temporal.py:545-547); semantics =
// SURVEY.md Appendix A.3 (stable descending order, topk cut, class-aware IoU > thresh in rank
// order, compaction, -1 fill).
//
// Pipeline inside the CTA:
//   1. gather <= 8 candidate lists (<= 1024 keys each) into registers
//   2. block_select_pivot -> compact the best <= SORTN keys to shared memory -> bitonic sort
//   3. n = min(k, #valid) ranks; gather box / class / area per rank (Source policy)
//   4. suppression bit-matrix  mask[r][p/32] bit p%32 = "r suppresses p" (p > r)
//        class-aware: secondary sort by (class, rank) so only same-class pairs are evaluated
//        class-agnostic / force_suppress: dense warp-ballot rows
//   5. one warp runs the sequential greedy scan over 32-rank blocks (diagonal words resolved
//      in-register, alive rows OR-ed into the running removed set)
//   6. survivors -> consecutive output rows (Sink policy)
#pragma once
#include "select.cuh"

namespace vd {

constexpr int kListCap = 1024;          // capacity of one candidate list
constexpr int kMaxLists = 8;            // lists merged per CTA
constexpr int kFinalThreads = 512;
constexpr int kFinalR = kListCap * kMaxLists / kFinalThreads;   // 16 keys / thread

struct NmsParams {
    float overlap_thresh;
    int k;              // min(topk, num_elem) <= VD_MAX_TOPK
    int sortn;          // power of two >= k, >= 512
    int class_aware;    // !force_suppress && id_index >= 0
    int max_out;        // rows to emit per image (post_nms for the fused head, k for box_nms)
    long long* dbg;     // optional per-phase clock64() stamps of block 0 (profiling aid), else null
};
#define VD_STAMP(P, i) do { if ((P).dbg && blockIdx.x == 0 && threadIdx.x == 0) (P).dbg[i] = clock64(); } while (0)

static inline int nms_sortn(int k) { return k <= 512 ? 512 : 1024; }
static inline int nms_words(int k) { return (k + 31) / 32; }
// dynamic shared memory of nms_final_kernel
static inline size_t nms_final_smem(int k) {
    size_t sortn = nms_sortn(k);
    size_t b = sortn * 8;                              // sorted keys
    b += (size_t)k * 16;                               // boxes (corner)
    b += (size_t)k * 4 * 2;                            // class, area
    b += (size_t)k * nms_words(k) * 4;                 // suppression matrix
    b += (size_t)sortn * 4;                            // (class,rank) secondary sort keys
    b += 64 * 4 + 64 * 4;                              // alive words + prefix
    b += 64 + 64;                                      // SelectScratch header + slack
    return b;
}

// bitonic sort of u32 (ascending), SN power of two
__device__ __forceinline__ void bitonic_sort_u32_asc(uint32_t* s, int SN) {
    for (int size = 2; size <= SN; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (SN >> 1); t += blockDim.x) {
                int i = 2 * t - (t & (stride - 1));
                int j = i + stride;
                bool asc = ((i & size) == 0);
                uint32_t a = s[i], b = s[j];
                if ((a > b) == asc) { s[i] = b; s[j] = a; }
            }
        }
    }
    __syncthreads();
}

// Bitonic sort with E = SN / blockDim.x elements per thread held in registers (thread t owns elements
// [t*E, t*E+E)): compare-exchange distances < E stay inside the thread, distances < 32*E use warp
// shuffles, larger ones go through shared memory -- 6 barrier rounds instead of 45 for SN = 512.
template <typename T, bool kDescending, int E>
__device__ __forceinline__ void bitonic_sort_regs(T* s, int SN) {
    const int tid = threadIdx.x;
    T v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = s[tid * E + e];
    auto cx = [&](T a, T o, int i, int stride, int size) -> T {       // value element i keeps after exchanging with i^stride
        const bool dir = ((i & size) == 0) == kDescending;             // this run sorts descending
        const bool lower = (i & stride) == 0;
        const T mx = a > o ? a : o, mn = a > o ? o : a;
        return (lower == dir) ? mx : mn;
    };
    for (int size = 2; size <= SN; size <<= 1) {
        int stride = size >> 1;
        for (; stride >= 32 * E; stride >>= 1) {
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; ++e) s[tid * E + e] = v[e];
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cx(v[e], s[(tid * E + e) ^ stride], tid * E + e, stride, size);
        }
        for (; stride >= E; stride >>= 1) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                T o;
                if constexpr (sizeof(T) == 8) o = (T)__shfl_xor_sync(0xffffffffu, (unsigned long long)v[e], stride / E);
                else o = (T)__shfl_xor_sync(0xffffffffu, (unsigned)v[e], stride / E);
                v[e] = cx(v[e], o, tid * E + e, stride, size);
            }
        }
#pragma unroll
        for (int st = E / 2; st >= 1; st >>= 1) {
            if (st < size) {
                T w[E];
#pragma unroll
                for (int e = 0; e < E; ++e) w[e] = cx(v[e], v[e ^ st], tid * E + e, st, size);
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = w[e];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) s[tid * E + e] = v[e];
    __syncthreads();
}
template <typename T, bool kDescending>
__device__ __forceinline__ void bitonic_sort_reg(T* s, int SN) { bitonic_sort_regs<T, kDescending, 1>(s, SN); }

// Block-wide sort of s[0..SN): register version when SN / blockDim.x is 1, 2, 4 or 8, else shared memory.
__device__ __forceinline__ void block_sort_u64_desc(uint64_t* s, int SN) {
    const int nt = (int)blockDim.x;
    if (SN == nt) bitonic_sort_regs<uint64_t, true, 1>(s, SN);
    else if (SN == 2 * nt) bitonic_sort_regs<uint64_t, true, 2>(s, SN);
    else if (SN == 4 * nt) bitonic_sort_regs<uint64_t, true, 4>(s, SN);
    else if (SN == 8 * nt) bitonic_sort_regs<uint64_t, true, 8>(s, SN);
    else bitonic_sort_desc(s, SN);
}
__device__ __forceinline__ void block_sort_u32_asc(uint32_t* s, int SN) {
    const int nt = (int)blockDim.x;
    if (SN == nt) bitonic_sort_regs<uint32_t, false, 1>(s, SN);
    else if (SN == 2 * nt) bitonic_sort_regs<uint32_t, false, 2>(s, SN);
    else if (SN == 4 * nt) bitonic_sort_regs<uint32_t, false, 4>(s, SN);
    else bitonic_sort_u32_asc(s, SN);
}

// Steps 3-6 on the sorted keys skeys[0..n): gather boxes, suppression matrix, greedy scan, emit.
template <class Source, class Sink>
__device__ void nms_tail(const int n, const int b, const NmsParams P, const Source& src, const Sink& sink,
                         uint64_t* skeys, float4* sbox, int* scls, float* sarea, uint32_t* smask,
                         uint32_t* skey2, uint32_t* salive, uint32_t* sprefix) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int k = P.k, SN = P.sortn, NW = (k + 31) >> 5;
    (void)k;
    VD_STAMP(P, 4);
    // ---- 3. per-rank box / class / area
    for (int r = tid; r < n; r += blockDim.x) {
        float4 bx; int c; float ar;
        src.load(b, key_row(skeys[r]), key_score(skeys[r]), bx, c, ar);
        sbox[r] = bx; scls[r] = c; sarea[r] = ar;
    }
    for (int i = tid; i < n * NW; i += blockDim.x) smask[i] = 0u;
    __syncthreads();

    VD_STAMP(P, 5);
    // ---- 4. suppression matrix
    if (P.class_aware) {
        // secondary key (class bucket, rank): same-class ranks become contiguous, rank ascending
        for (int i = tid; i < SN; i += blockDim.x)
            skey2[i] = (i < n) ? (((uint32_t)scls[i] << 10) | (uint32_t)i) : 0xffffffffu;
        __syncthreads();
        block_sort_u32_asc(skey2, SN);
        VD_STAMP(P, 6);
        // 4 threads per rank walk the rank's bucket tail with stride 4 (buckets are ~n/C long)
        for (int i = tid >> 2; i < n; i += blockDim.x >> 2) {
            uint32_t ki = skey2[i];
            int r = (int)(ki & 1023u); uint32_t ci = ki >> 10;
            float4 br = sbox[r]; float ar = sarea[r]; const int cr = scls[r];
            for (int j = i + 1 + (tid & 3); j < n; j += 4) {
                uint32_t kj = skey2[j];
                if ((kj >> 10) != ci) break;
                int p = (int)(kj & 1023u);             // p > r (rank ascending inside a bucket)
                if (scls[p] != cr) continue;           // bucket = low 22 bits of the id; compare the full int
                if (vd_iou_gt(br, ar, sbox[p], sarea[p], P.overlap_thresh))
                    atomicOr(&smask[r * NW + (p >> 5)], 1u << (p & 31));
            }
        }
    } else {
        for (int r = warp; r < n; r += nwarps) {
            float4 br = sbox[r]; float ar = sarea[r];
            for (int w = r >> 5; w < NW; ++w) {
                int p = (w << 5) + lane;
                bool s = false;
                if (p > r && p < n) s = vd_iou_gt(br, ar, sbox[p], sarea[p], P.overlap_thresh);
                unsigned m = __ballot_sync(0xffffffffu, s);
                if (lane == 0) smask[r * NW + w] = m;
            }
        }
    }
    __syncthreads();

    VD_STAMP(P, 7);
    // ---- 5. greedy scan (warp 0): lane l owns removed-bits of ranks 32l .. 32l+31
    if (warp == 0) {
        uint32_t removed = 0u;
        int kept_total = 0;
        for (int w = 0; w < NW; ++w) {
            const int base = w << 5;
            uint32_t cur = __shfl_sync(0xffffffffu, removed, w);
            const int rr = base + lane;
            uint32_t diag = (rr < n) ? smask[rr * NW + w] : 0u;
            unsigned nz = __ballot_sync(0xffffffffu, diag != 0u);
            while (nz) {                               // rows of this block that suppress inside it
                int j = __ffs(nz) - 1; nz &= nz - 1;
                uint32_t dj = __shfl_sync(0xffffffffu, diag, j);
                if (!((cur >> j) & 1u)) cur |= dj;
            }
            uint32_t validbits = (n - base >= 32) ? 0xffffffffu : (n > base ? ((1u << (n - base)) - 1u) : 0u);   // blocks past n: no ranks
            uint32_t alive = ~cur & validbits;
            if (lane == w) removed = cur;
            if (lane == 0) { salive[w] = alive; sprefix[w] = (uint32_t)kept_total; }
            kept_total += __popc(alive);
            if (kept_total >= P.max_out) {             // later ranks cannot reach the output
                for (int w2 = w + 1 + lane; w2 < NW; w2 += 32) { salive[w2] = 0u; sprefix[w2] = (uint32_t)kept_total; }
                break;
            }
            // alive rows suppress later blocks: 32 independent (warp-uniformly predicated) row loads
            if (lane > w && lane < NW) {
                const uint32_t* mrow = smask + base * NW + lane;
                uint32_t acc0 = 0u, acc1 = 0u;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    if ((alive >> j) & 1u) acc0 |= mrow[j * NW];
                    if ((alive >> (j + 1)) & 1u) acc1 |= mrow[(j + 1) * NW];
                }
                removed |= acc0 | acc1;
            }
        }
        if (lane == 0) sprefix[63] = (uint32_t)kept_total;
    }
    __syncthreads();

    VD_STAMP(P, 8);
    // ---- 6. emit survivors
    const int kept_total = (int)sprefix[63];
    for (int r = tid; r < n; r += blockDim.x) {
        uint32_t aw = salive[r >> 5];
        if ((aw >> (r & 31)) & 1u) {
            int pos = (int)sprefix[r >> 5] + __popc(aw & ((1u << (r & 31)) - 1u));
            if (pos < P.max_out) sink.emit(b, pos, key_row(skeys[r]), key_score(skeys[r]), sbox[r], scls[r]);
        }
    }
    sink.finish(b, kept_total < P.max_out ? kept_total : P.max_out);
    VD_STAMP(P, 9);
}
// Source: gives box (corner), integer class and area for (image b, row).  Sink: writes output.
template <class Source, class Sink>
__device__ void nms_final_body(const uint64_t* __restrict__ lists, const uint32_t* __restrict__ counts,
                               int n_lists, int b, NmsParams P, const Source& src, const Sink& sink,
                               unsigned char* smem_raw) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int k = P.k, SN = P.sortn, NW = (k + 31) >> 5;
    // ---- carve shared memory
    SelectScratch* scr = reinterpret_cast<SelectScratch*>(smem_raw);          // 8-byte members: keep it first
    uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + 64);
    float4* sbox = reinterpret_cast<float4*>(skeys + SN);
    int* scls = reinterpret_cast<int*>(sbox + k);
    float* sarea = reinterpret_cast<float*>(scls + k);
    uint32_t* smask = reinterpret_cast<uint32_t*>(sarea + k);
    uint32_t* skey2 = smask + (size_t)k * NW;
    uint32_t* salive = skey2 + SN;
    uint32_t* sprefix = salive + 64;
    static_assert(sizeof(SelectScratch) <= 64, "scratch header");

    select_scratch_init(scr);
    for (int i = tid; i < SN; i += blockDim.x) skeys[i] = 0ull;

    // ---- 1. gather lists into registers
    uint64_t keys[kFinalR];
#pragma unroll
    for (int r = 0; r < kFinalR; ++r) {
        int slot = r * kFinalThreads + tid;            // list = slot / kListCap
        int l = slot / kListCap, j = slot % kListCap;
        uint64_t v = 0ull;
        if (l < n_lists) {
            uint32_t c = counts[l]; c = c > (uint32_t)kListCap ? (uint32_t)kListCap : c;
            if ((uint32_t)j < c) v = lists[(size_t)l * kListCap + j];
        }
        keys[r] = v;
    }
    // ---- 2. select + sort
    int it = 0; uint32_t nsel = 0;
    uint64_t piv = block_select_pivot<kFinalR>(keys, (uint32_t)k, (uint32_t)SN, 0ull, 0ull, scr, it, &nsel);
    block_compact<kFinalR>(keys, piv, skeys, (uint32_t)SN, &scr->out_count);
    __syncthreads();
    block_sort_u64_desc(skeys, SN);                    // starts and ends with __syncthreads
    const int n = (int)(nsel < (uint32_t)k ? nsel : (uint32_t)k);
    nms_tail(n, b, P, src, sink, skeys, sbox, scls, sarea, smask, skey2, salive, sprefix);
}


// Steps 3-6 of the per-image NMS as used by the fused head (head.cu) and by box_nms with more than VD_MAX_TOPK candidates (nms.cu): class-aware greedy NMS as a WAVEFRONT over 32-rank chunks, with no
// n x n matrix and no class sort.  Only KEPT boxes ever suppress and the scan stops once max_out boxes are kept, so a
// chunk's members are tested just in time against the (< max_out + 32) boxes kept so far:
//   round c, test phase (all warps):  lane = member of chunk c; warp w tests it against kept boxes w, w+8, ... and
//       against rows 4w..4w+3 of the chunk itself (one ballot per row = the chunk's 32x32 "diagonal" block)
//   round c, resolve phase (warp 0):  greedy rank-order resolution of the chunk in registers, survivors are emitted
//       straight to the output and appended to the kept list
// Semantics = nms_core.cuh::nms_tail (rank order, IoU > thresh, same class, first max_out survivors), same IoU arithmetic.
constexpr int kNmsThreads = 256;
struct WaveShared { uint32_t kept_total; uint32_t sup; uint32_t diag[32]; };
template <class Source, class Sink>
__device__ void nms_tail_wave(const int n, const int f, const NmsParams P, const Source& src, const Sink& sink,
                              const uint64_t* skeys, float4* sbox, int* scls, float* sarea,
                              float4* skbox, float* skarea, int* skcls, WaveShared* wsh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int NB = (n + 31) >> 5;
    const float thr = P.overlap_thresh;
    VD_STAMP(P, 4);
    for (int r = tid; r < n; r += blockDim.x) {
        float4 bx; int c; float ar;
        src.load(f, key_row(skeys[r]), key_score(skeys[r]), bx, c, ar);
        sbox[r] = bx; scls[r] = c; sarea[r] = ar;
    }
    if (tid == 0) { wsh->kept_total = 0u; wsh->sup = 0u; }
    __syncthreads();
    VD_STAMP(P, 5);
    uint32_t kept_total = 0u;
    for (int c = 0; c < NB; ++c) {
        // ---- test phase
        const int pm = 32 * c + lane;
        const bool valid = pm < n;
        float4 bm = make_float4(0.f, 0.f, 0.f, 0.f); float am_ = 0.f; int cm = -2;
        if (valid) { bm = sbox[pm]; am_ = sarea[pm]; cm = scls[pm]; }
        {
            bool sup = false;
            for (uint32_t i0 = (uint32_t)warp; i0 < kept_total; i0 += 4u * (uint32_t)nwarps) {     // kept list is padded with never-matching entries
                bool sp[4], am[4]; bool any_amb = false;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t i = i0 + (uint32_t)(u * nwarps);
                    sp[u] = vd_iou_gt_fast(skbox[i], skarea[i], bm, am_, thr, am[u]);
                    const bool same = skcls[i] == cm;
                    sp[u] &= same; am[u] &= same; any_amb |= am[u];
                }
                if (__builtin_expect(any_amb, 0)) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const uint32_t i = i0 + (uint32_t)(u * nwarps); if (am[u]) sp[u] = vd_iou_gt(skbox[i], skarea[i], bm, am_, thr); }
                }
                sup |= sp[0] | sp[1] | sp[2] | sp[3];
            }
            const unsigned sw = __ballot_sync(0xffffffffu, sup);
            if (lane == 0 && sw) atomicOr(&wsh->sup, sw);
            // rows 4w .. 4w+3 of the chunk's own block (row = earlier rank, lane = later rank)
            bool sp[4], am[4]; bool any_amb = false;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = 4 * warp + u, pr = min(32 * c + rr, n - 1);
                sp[u] = vd_iou_gt_fast(sbox[pr], sarea[pr], bm, am_, thr, am[u]);
                const bool rel = (lane > rr) & (scls[pr] == cm) & (32 * c + rr < n);
                sp[u] &= rel; am[u] &= rel; any_amb |= am[u];
            }
            if (__builtin_expect(any_amb, 0)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int pr = min(32 * c + 4 * warp + u, n - 1); if (am[u]) sp[u] = vd_iou_gt(sbox[pr], sarea[pr], bm, am_, thr); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned bal = __ballot_sync(0xffffffffu, sp[u]);
                if (lane == 0) wsh->diag[4 * warp + u] = bal;
            }
        }
        __syncthreads();
        // ---- resolve phase
        if (warp == 0) {
            uint32_t cur = wsh->sup | (valid ? 0u : 0u);
            cur |= ~__ballot_sync(0xffffffffu, valid);                 // ranks past n never survive
            const uint32_t diag = wsh->diag[lane];
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
                uint32_t d[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) d[j] = __shfl_sync(0xffffffffu, diag, j0 + j);
#pragma unroll
                for (int j = 0; j < 8; ++j) if (!((cur >> (j0 + j)) & 1u)) cur |= d[j];
            }
            const uint32_t alive = ~cur;
            const uint32_t cnt = (uint32_t)__popc(alive);
            const uint32_t slot = kept_total + (uint32_t)__popc(alive & ((1u << lane) - 1u));
            if ((alive >> lane) & 1u) {
                if (slot < (uint32_t)P.max_out) sink.emit(f, (int)slot, key_row(skeys[pm]), key_score(skeys[pm]), bm, cm);
                VD_DEV_CHECK(slot < (uint32_t)((P.max_out < P.k ? P.max_out : P.k) + 32));
                skbox[slot] = bm; skarea[slot] = am_; skcls[slot] = cm;
            }
            // pad the list to the test loop's stride with entries that match no class
            const uint32_t padded = (kept_total + cnt + 4u * (uint32_t)nwarps);
            for (uint32_t i = kept_total + cnt + (uint32_t)lane; i < padded; i += 32u) skcls[i] = -3;
            if (lane == 0) { wsh->kept_total = kept_total + cnt; wsh->sup = 0u; }
        }
        __syncthreads();
        kept_total = wsh->kept_total;
        if (kept_total >= (uint32_t)P.max_out) break;
    }
    VD_STAMP(P, 8);
    sink.finish(f, (int)kept_total < P.max_out ? (int)kept_total : P.max_out);
    VD_STAMP(P, 9);
}

// Intermediate level: merge <= 8 lists into one list holding a superset (<= kListCap) of their
// joint top-k.  grid (n_groups, num_batch).
static __global__ void __launch_bounds__(kFinalThreads, 1)
nms_merge_kernel(const uint64_t* __restrict__ lists_in, const uint32_t* __restrict__ counts_in,
                 int n_lists_in, uint64_t* __restrict__ lists_out, uint32_t* __restrict__ counts_out,
                 int n_lists_out, int k) {
    __shared__ SelectScratch scr;
    __shared__ uint64_t stage[kListCap];
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    select_scratch_init(&scr);
    const uint64_t* lin = lists_in + (size_t)b * n_lists_in * kListCap;
    const uint32_t* cin = counts_in + (size_t)b * n_lists_in;
    uint64_t keys[kFinalR];
#pragma unroll
    for (int r = 0; r < kFinalR; ++r) {
        int slot = r * kFinalThreads + tid;
        int l = g * kMaxLists + slot / kListCap, j = slot % kListCap;
        uint64_t v = 0ull;
        if (l < n_lists_in) {
            uint32_t c = cin[l]; c = c > (uint32_t)kListCap ? (uint32_t)kListCap : c;
            if ((uint32_t)j < c) v = lin[(size_t)l * kListCap + j];
        }
        keys[r] = v;
    }
    int it = 0; uint32_t nsel = 0;
    uint64_t piv = block_select_pivot<kFinalR>(keys, (uint32_t)k, (uint32_t)kListCap, 0ull, 0ull, &scr, it, &nsel);
    block_compact<kFinalR>(keys, piv, stage, (uint32_t)kListCap, &scr.out_count);
    __syncthreads();
    uint64_t* lout = lists_out + ((size_t)b * n_lists_out + g) * kListCap;
    for (int i = tid; i < (int)nsel; i += blockDim.x) lout[i] = stage[i];
    if (tid == 0) counts_out[(size_t)b * n_lists_out + g] = nsel;
}

}  // namespace vd
