// Block-level exact top-k machinery shared by box_nms and the fused head:
//   * block_sum            -- one-barrier block reduction (3 rotating shared counters)
//   * block_select_pivot   -- finds a pivot key P with  min(k,n) <= #{key >= P} <= cap  by
//                             guess -> gallop -> bisection on the 64-bit key (keys are unique)
//   * block_compact        -- warp-aggregated compaction of the keys >= P into shared memory
//   * bitonic_sort_desc    -- shared-memory bitonic sort, descending
// Exactness argument: an element of the global top-k has < k elements ahead of it globally, hence
// < k ahead of it inside any subset that contains it, so "filter each subset to a superset of its
// own top-k, then re-select" never loses a member of the global top-k.
#pragma once
#include "common.cuh"

namespace vd {

struct SelectScratch {
    uint32_t cnt[3];     // rotating counters for block_sum (must start at 0)
    uint32_t out_count;  // compaction cursor
    uint64_t red[2];     // min / max reduction slots
};

__device__ __forceinline__ void select_scratch_init(SelectScratch* s) {
    if (threadIdx.x == 0) {
        s->cnt[0] = s->cnt[1] = s->cnt[2] = 0; s->out_count = 0;
        s->red[0] = ~0ull; s->red[1] = 0ull;
    }
    __syncthreads();
}

// Sum of v over the block; every thread gets the result.  `it` is a per-thread iteration counter
// that must advance identically in all threads.
__device__ __forceinline__ uint32_t block_sum(uint32_t v, SelectScratch* s, int& it) {
    uint32_t w = __reduce_add_sync(0xffffffffu, v);
    const int slot = it % 3;
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s->cnt[slot], w);
    if (threadIdx.x == 0) s->cnt[(it + 1) % 3] = 0;
    __syncthreads();
    uint32_t r = s->cnt[slot];
    ++it;
    return r;
}

template <int R>
__device__ __forceinline__ uint32_t count_ge(const uint64_t (&keys)[R], uint64_t piv) {
    uint32_t c = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) c += (keys[r] >= piv) ? 1u : 0u;
    return c;
}

// Returns pivot P (>= 1).  On return *n_sel = #{key >= P}.  Requires cap >= k >= 1, unique
// non-zero keys (0 = invalid slot).  `guess` = warm start (0 = none), `gallop0` = first gallop step.
template <int R>
__device__ uint64_t block_select_pivot(const uint64_t (&keys)[R], uint32_t k, uint32_t cap,
                                       uint64_t guess, uint64_t gallop0, SelectScratch* s,
                                       int& it, uint32_t* n_sel) {
    uint32_t n = block_sum(count_ge<R>(keys, 1ull), s, it);
    if (n <= cap) { *n_sel = n; return 1ull; }
    // invariant: count(>= lo) > cap  and  count(>= hi) < k
    uint64_t lo = 1ull, hi = ~0ull;
    bool have_lo = false, have_hi = false;
    uint64_t mid;
    if (guess > 1ull) {
        mid = guess;
        uint64_t step = gallop0 ? gallop0 : (1ull << 49);
        for (int g = 0; g < 64; ++g) {                       // gallop until bracketed
            uint32_t t = block_sum(count_ge<R>(keys, mid), s, it);
            if (t > cap) { lo = mid; have_lo = true; if (have_hi) break;
                           uint64_t nm = mid + step; if (nm < mid) { break; } mid = nm; }
            else if (t < k) { hi = mid; have_hi = true; if (have_lo) break;
                              if (mid <= step + 1ull) { break; } mid -= step; }
            else { *n_sel = t; return mid; }
            step <<= 1; if (step == 0) break;
        }
    }
    if (!have_lo || !have_hi) {
        // cold bracket from the block's min / max valid key
        uint64_t mn = ~0ull, mx = 0ull;
#pragma unroll
        for (int r = 0; r < R; ++r) if (keys[r]) { mn = keys[r] < mn ? keys[r] : mn; mx = keys[r] > mx ? keys[r] : mx; }
        for (int o = 16; o > 0; o >>= 1) {
            uint64_t a = __shfl_xor_sync(0xffffffffu, mn, o); mn = a < mn ? a : mn;
            uint64_t b = __shfl_xor_sync(0xffffffffu, mx, o); mx = b > mx ? b : mx;
        }
        if ((threadIdx.x & 31) == 0) { atomicMin((unsigned long long*)&s->red[0], (unsigned long long)mn);
                                       atomicMax((unsigned long long*)&s->red[1], (unsigned long long)mx); }
        __syncthreads();
        mn = s->red[0]; mx = s->red[1];
        __syncthreads();
        if (threadIdx.x == 0) { s->red[0] = ~0ull; s->red[1] = 0ull; }
        if (!have_lo) lo = mn;            // count(>= min) = n > cap
        if (!have_hi) hi = mx + 1ull;     // count(>  max) = 0 < k
    }
    for (int g = 0; g < 80; ++g) {
        mid = lo + ((hi - lo) >> 1);
        uint32_t t = block_sum(count_ge<R>(keys, mid), s, it);
        if (t > cap) lo = mid;
        else if (t < k) hi = mid;
        else { *n_sel = t; return mid; }
    }
    // unreachable for unique keys; keep everything above lo (superset, still exact after re-select)
    *n_sel = block_sum(count_ge<R>(keys, lo), s, it);
    return lo;
}

// Append every key >= piv to dst[] (arbitrary order); cursor is a shared counter.
template <int R>
__device__ __forceinline__ void block_compact(const uint64_t (&keys)[R], uint64_t piv, uint64_t* dst,
                                              uint32_t dst_cap, uint32_t* cursor) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        bool p = keys[r] >= piv;
        unsigned m = __ballot_sync(0xffffffffu, p);
        if (m) {
            int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            if (p && pos < dst_cap) dst[pos] = keys[r];
        }
    }
}

// In-place descending bitonic sort of s[0..SN) (SN power of two), all threads of the block.
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* s, int SN) {
    for (int size = 2; size <= SN; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (SN >> 1); t += blockDim.x) {
                int i = 2 * t - (t & (stride - 1));
                int j = i + stride;
                bool desc = ((i & size) == 0);
                uint64_t a = s[i], b = s[j];
                if ((a < b) == desc) { s[i] = b; s[j] = a; }
            }
        }
    }
    __syncthreads();
}

// IoU test of MXNet's nms_impl with its exact fp32 operation order:
//   iou = inter / (area_r + area_p - inter);  suppressed iff iou > thresh.
__device__ __forceinline__ float vd_intersect_1d(float a1, float a2, float b1, float b2) {
    float left = a1 > b1 ? a1 : b1;
    float right = a2 < b2 ? a2 : b2;
    float w = __fsub_rn(right, left);
    return w > 0.0f ? w : 0.0f;
}
// MXNet BoxArea(): width * height, 0 when either extent is negative (oracle/ASSUMPTIONS.md A3)
__device__ __forceinline__ float vd_box_area(float w, float h) { return (w < 0.0f || h < 0.0f) ? 0.0f : __fmul_rn(w, h); }

__device__ __forceinline__ bool vd_iou_gt(float4 r, float area_r, float4 p, float area_p, float thresh) {
    float inter = __fmul_rn(vd_intersect_1d(r.x, r.z, p.x, p.z), vd_intersect_1d(r.y, r.w, p.y, p.w));
    if (inter == 0.0f && thresh >= 0.0f) return false;     // 0/u is 0, -0 or NaN: never > thresh >= 0
    const float uni = __fsub_rn(__fadd_rn(area_r, area_p), inter);
    // Decide without the IEEE division when inter is clearly on one side of thresh*union: for
    // positive finite uni,  inter/uni > thresh  <=>  inter > thresh*uni  exactly in real arithmetic;
    // the two fp32 roundings (product, quotient) can only matter within a few ulp of equality.
    if (uni > 0.0f && thresh >= 0.0f && inter < 3.0e38f && uni < 3.0e38f) {
        const float tu = __fmul_rn(thresh, uni);
        const float margin = __fmul_rn(tu, 4.0e-7f) + 1.0e-37f;
        if (inter > tu + margin) return true;
        if (inter < tu - margin) return false;
    }
    float iou = __fdiv_rn(inter, uni);
    return iou > thresh;
}

// Straight-line form of vd_iou_gt for unrolled loops (few warps per SM: independent tests must interleave, so no
// branches inside): returns the decision when inter is clearly on one side of thresh*union (same margin as vd_iou_gt),
// and sets `amb` for the few-ulp band around equality and for degenerate unions -- the caller resolves those with
// vd_iou_gt (the IEEE division) outside its unrolled group.  Decisions are identical to vd_iou_gt by construction.
__device__ __forceinline__ bool vd_iou_gt_fast(float4 r, float area_r, float4 p, float area_p, float thresh, bool& amb) {
    const float inter = __fmul_rn(vd_intersect_1d(r.x, r.z, p.x, p.z), vd_intersect_1d(r.y, r.w, p.y, p.w));
    const float uni = __fsub_rn(__fadd_rn(area_r, area_p), inter);
    const float tu = __fmul_rn(thresh, uni);
    const float margin = __fmul_rn(tu, 4.0e-7f) + 1.0e-37f;
    const bool common = (uni > 0.0f) & (thresh >= 0.0f) & (inter < 3.0e38f) & (uni < 3.0e38f);
    const bool zero = (inter == 0.0f) & (thresh >= 0.0f);
    const bool yes = common & (inter > tu + margin) & !zero;
    const bool no = zero | (common & (inter < tu - margin));
    amb = !(yes | no);
    return yes;
}

}  // namespace vd
