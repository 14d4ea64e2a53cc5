// temporal_head_fused_kernel -- cfg 4 in ONE kernel, all scales in one launch (SURVEY section 7, K4 -> K1):
//   temporal (3,1,1) tip cell (Conv3D + BN + LeakyReLU, layers.py:82-89; yolo3_temporal.py:226-227)
//   -> 1x1 prediction conv (yolo3.py:62,157) -> YOLOOutputV3 decode (yolo3.py:158-199) -> speculative candidate filter
// The tip tile never leaves the SM: the tip GEMM's accumulator (TMEM) goes through BN / LeakyReLU / bf16 rounding into shared
// memory in the K-major 128-byte-swizzled layout of a UMMA A operand, and the prediction GEMM multiplies it from there.
// Included by head.cu (uses HeadGeom / kSpecCap and the decode math of common.cuh); results are bit-identical to the unfused
// chain temporal_conv_pair_kernel -> head_kernel<EPI_SPEC> (same bf16 tip bits, same K order of the prediction GEMM).
//
// CTA pair (tcgen05 cta_group::2), persistent; an item = two consecutive 128-row tiles of the flattened (window, T*HW) row axis of
// one scale, all its channel chunks of 256; the host deals the items of all scales to the pairs (tfused_schedule):
//   warp 0 (both CTAs)   TMA producer.  Ring entries in consumption order: per k-block the CTA's A tile [128 rows x 64 ch] of the
//                        shifted frame + half of the tap's weight tile [128 x 64]; behind the k-blocks of chunk c + 1 one entry with
//                        the CTA's half of chunk c's prediction weights [NPAD/2 x 256 ch]
//   warp 1 (leader CTA)  MMA issuer: tip GEMM M256 x N256 x K(3 taps x Cin) of chunk c + 1 into TMEM cols [0,256) (single buffer),
//                        commit, then the prediction GEMM M256 x NPAD x K256 of chunk c from the staged tip chunk into one of two
//                        prediction accumulators (TMEM cols [256,384) / [384,512), accumulating over an item's chunks): the tensor
//                        pipe runs it while the epilogue warps read the tip accumulator back
//   warps 2-9            per chunk: tcgen05.ld of the warp's 128 columns -> accumulator handed back -> BN (shared memory) -> LeakyReLU
//                        -> bf16 -> st.shared (swizzled) -> fence.proxy.async -> arrive; then, one chunk late, the decode + candidate
//                        filter of the item whose last prediction GEMM has just run (the EPI_SPEC epilogue of head_kernel; a tile
//                        spans up to two frames here, so frame / cell are per lane and candidates go to their frame's list with one
//                        atomic each -- ~0.3 % of the class logits pass)
// The tip accumulator is single-buffered (TMEM: 256 + 2 x 128 columns); the tip GEMM of these shapes is bound by the operand traffic
// L2 -> SM rather than by the tensor pipe (same time at 1.9 and 1.5 GHz SM clock), so the ring keeps filling during the read-back.
// Measured step by step in DESIGN.md section 4.5.
#pragma once

namespace vd {

constexpr int F_THREADS = 320;
constexpr int F_BLOCK_M = 128;
constexpr int F_BLOCK_K = 64;
constexpr int F_NT = 256;                    // tip channels per chunk = K of one prediction-GEMM step
constexpr int F_MAX_CLUSTERS = 80;           // CTA pairs of one launch (sm_count / 2 <= 80)

// Class-bias tables of spec_decode_lane in shared memory, per scale: [3 anchors][CPA chunks][CH4] biases (chunks padded for 128-bit
// loads), then [3][CPA] = the largest bias of each chunk (the class loop's first test compares the raw logits against
// threshold - max bias: no add and no bias load unless a logit of the chunk can pass).
template <int C> struct SpecTables {
    static constexpr int CPA = (C + 15) / 16;
    static constexpr int CH = (C + CPA - 1) / CPA;
    static constexpr int CH4 = (CH + 3) / 4 * 4;
    static constexpr int NCH = 3 * CPA;
    static constexpr int BLK = (NCH * CH4 + NCH + 3) / 4 * 4;          // floats per scale
};
template <int C>
__device__ __forceinline__ void spec_stage_tables(float* scbias, const float* bias, int s, int tid, int nthreads) {
    using T = SpecTables<C>;
    constexpr int P = 5 + C;
    float* blk = scbias + s * T::BLK;
    for (int i = tid; i < T::NCH * T::CH4; i += nthreads) {
        const int a = i / (T::CPA * T::CH4), cc = (i / T::CH4) % T::CPA, ci = i % T::CH4;
        const int c = cc * T::CH + ci;
        blk[i] = (bias && ci < T::CH && c < C) ? bias[a * P + 5 + c] : 0.0f;
    }
    for (int j = tid; j < T::NCH; j += nthreads) {
        const int a = j / T::CPA, cc = j - a * T::CPA;
        float m = 0.0f;                                                // padding classes carry bias 0
        for (int ci = 0; ci < T::CH; ++ci) { const int c = cc * T::CH + ci; if (bias && c < C) m = fmaxf(m, bias[a * P + 5 + c]); }
        blk[T::NCH * T::CH4 + j] = m;
    }
}

// One launch covers every scale: an s32 item (4 chunks of K = 3072) costs ~14x an s8 item, and 224 of them over 74 CTA pairs were
// 3.03 waves (a quarter of the launch lost to the last one).  The host deals the items of all scales to the pairs, largest first
// (longest-processing-time greedy on a cost model), as one contiguous range per (scale, pair): beg[s][c] .. beg[s][c + 1].
struct FusedScale {
    int HW, Cin, rows, m_tiles, n_chunks;    // pixels per frame, channels, rows = T*HW, 128-row tiles per window, Cin / 256
    const float* scale; const float* shift;  // folded BN of the tip cell
    const float* bias;                       // prediction bias (3*(5+C)) or null
};
struct FusedParams {
    int B, T, num_scales;                    // windows, frames per window
    float slope;
    FusedScale sc[VD_MAX_SCALES];
    unsigned short beg[VD_MAX_SCALES][F_MAX_CLUSTERS + 1];
    // scales whose +-HW neighbour rows lie many tiles apart (s8: 21) first run `strided[s]` rounds of item = round * pairs + pair: tile m
    // and the tiles its outer taps read (m +- HW/128) are then in flight on different pairs at the same time and share L2 (with
    // contiguous ranges alone they are ~10 items = ~100 us apart on ONE pair: DRAM reads doubled); beg[][] covers what is left
    int strided[VD_MAX_SCALES], pairs;
    HeadGeom g;
    int c_valid;                             // classes actually present (<= C)
    float valid_thresh;
    float4* boxes; uint64_t* spec_lists; uint32_t* spec_cnt; const uint32_t* spec_tau;
    const unsigned int* tile_counter; unsigned int ws_magic;
    int frames;
    int dbg;                                 // profiling aid (VD_TFUSED_DBG): 1 = skip the decode / filter epilogue
    long long* stamps;                       // profiling aid (VD_TFUSED_STAMPS): clock64 per chunk of cluster 0's leader CTA, [chunk][16]
};
struct FusedMaps { CUtensorMap x[VD_MAX_SCALES], w[VD_MAX_SCALES], wp[VD_MAX_SCALES]; };

template <int C, int NPAD> struct FusedCfg {
    static constexpr int A_BYTES = F_BLOCK_M * F_BLOCK_K * 2;
    static constexpr int B_BYTES = (F_NT / 2) * F_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = 4;
    static constexpr int STG_TILE = F_BLOCK_M * F_BLOCK_K * 2;            // one k-block of the staged tip chunk
    static constexpr int STG_BYTES = (F_NT / F_BLOCK_K) * STG_TILE;
    static constexpr int WP_ROWS = NPAD / 2;                              // this CTA's half of the prediction weights
    static constexpr int WP_TILE = WP_ROWS * F_BLOCK_K * 2;
    static constexpr int WP_BYTES = (F_NT / F_BLOCK_K) * WP_TILE;          // one chunk's prediction weights travel through ONE ring stage
    static constexpr int CPA = (C + 15) / 16;
    static constexpr int CH = (C + CPA - 1) / CPA;
    static constexpr int CH4 = (CH + 3) / 4 * 4;
    static constexpr int CBIAS_BYTES = VD_MAX_SCALES * SpecTables<C>::BLK * 4;
    static constexpr int BIAS_BYTES = VD_MAX_SCALES * NPAD * 4;
    static constexpr int BN_BYTES = VD_MAX_SCALES * 2 * 1024 * 4;          // folded BN scale / shift of every scale (Cin <= 1024)
    static constexpr int SH_BYTES = 1024;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BN_BYTES + CBIAS_BYTES + BIAS_BYTES + SH_BYTES + 1024;
    static_assert(WP_ROWS % 8 == 0 && NPAD % 16 == 0 && NPAD <= 128, "prediction width");
    static_assert(3 * (5 + C) <= NPAD, "NPAD");
    static_assert(SMEM_BYTES <= 227 * 1024 && WP_BYTES <= STAGE_BYTES, "shared memory");
};

struct FusedShared {
    uint64_t full[4], empty[4];
    uint64_t tip_full, tip_empty, stg_full, stg_empty;
    uint64_t pred_full[2], pred_empty[2];
    uint32_t tmem_base;
};

// Where the decode + speculative candidate filter of one accumulator row puts its results (head workspace, see head.cu).
struct SpecOut {
    int rows_total, anc_total;               // rows of a frame's (rows, 6) tensor / anchor slots of a frame, over all scales
    int c_valid; float valid_thresh;
    float4* boxes; uint64_t* spec_lists; uint32_t* spec_cnt; const uint32_t* spec_tau;
    int frames;
};

// Decode + speculative candidate filter of ONE pixel's prediction logits (the EPI_SPEC epilogue of head_kernel with a per-lane frame):
// `tb` = TMEM address of the lane's accumulator row (3 anchors x (5 + C) columns), `half` / `nhalf` = which share of the 3*CPA class
// chunks this warp takes (the warps that share a TMEM lane quarter split them), (f, cell) = frame and pixel of the lane (inb = false:
// padding row, nothing is emitted), row_base_s / anc_base_s = the scale's first row / anchor slot.  sbias_s [NPAD] / scbias_s (SpecTables block) = the scale's biases in shared memory.
// Every candidate whose score can reach the frame slot's threshold tau is scored and appended to its FRAME's list with one atomic
// (~0.3 % of the class logits pass); the raw box records of the emitting (pixel, anchor) pairs go to the workspace.
template <int C>
__device__ __forceinline__ void spec_decode_lane(const uint32_t tb, const int half, const int nhalf, const bool inb, const int f, const int cell, const int row_base_s, const int anc_base_s, const int HW,
                                                 const float* sbias_s, const float* scbias_s, const bool ws_ok, const uint32_t tau_hint, const SpecOut& o) {
    constexpr int P = 5 + C;
    constexpr int CPA_ = (C + 15) / 16, CH_ = (C + CPA_ - 1) / CPA_, CH4_ = (CH_ + 3) / 4 * 4;
        uint32_t spec_tb;
        {
            const uint32_t floor_b = o.valid_thresh > 0.0f ? __float_as_uint(o.valid_thresh) : 0u;
            const uint32_t hint = ws_ok ? tau_hint : 0u;          // the frame slot's threshold, loaded by the caller BEFORE it waits for the accumulator
            spec_tb = hint > floor_b ? hint : floor_b;
            if (spec_tb > 0x3f800001u) spec_tb = 0x3f800001u;
            if (!ws_ok) spec_tb = 0x3f800001u;                            // foreign workspace: emit nothing, the NMS kernel fails every frame
        }
        constexpr int CH = CH_, CPA = CPA_, CH4 = CH4_, REM = C - (CPA - 1) * CH;
        constexpr int NCH = 3 * CPA;
        uint32_t rc[2][CH];
        auto issue_j = [&](const int j, uint32_t* dst) {
            const int a = j / CPA, cx = j - a * CPA;
            const uint32_t col = tb + (uint32_t)(a * P + 5 + cx * CH);
            if (REM != CH && cx == CPA - 1) tc::tmem_ld<REM>(col, dst); else tc::tmem_ld<CH>(col, dst);
        };
        float conf[3];
        {
            uint32_t rb[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) tc::tmem_ld1(tb + (uint32_t)(a * P + 4), rb + a);
            if (half < NCH) issue_j(half, rc[0]);                         // the warp's first class chunk rides along with the objectness columns
            tc::tmem_ld_wait();
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                conf[a] = vd_sigmoid(__uint_as_float(rb[a]) + sbias_s[a * P + 4]);
                if (!inb) conf[a] = __uint_as_float(0x7fc00000u);          // NaN: no score of a padding row passes `> valid_thresh`
            }
        }
        const float vth = o.valid_thresh;
        float ell[3];
        {
            const float t = __fmul_rn(__uint_as_float(spec_tb), 0.999969482421875f);   // 1 - 2^-15
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float rr = __fmul_rn(t, vd_rcp(conf[a]));
                const float l = __fmul_rn(__fsub_rn(vd_lg2(rr), vd_lg2(__fsub_rn(1.0f, rr))), 0.6931471805599453f);
                float e = (rr < 1.0f) ? l : __uint_as_float(0x7f800000u);
                if (spec_tb == 0u) e = __uint_as_float(0xff800000u);
                if (!(conf[a] == conf[a])) e = __uint_as_float(0x7f800000u);
                if (spec_tb >= 0x3f800001u) e = __uint_as_float(0x7f800000u);
                ell[a] = e;
            }
        }
        const uint32_t HW3 = (uint32_t)HW * 3u;
        const uint32_t row0 = (uint32_t)row_base_s + (uint32_t)cell * 3u;       // + c*HW3 + a
        uint64_t* fl = o.spec_lists + (size_t)f * kSpecCap;
        uint32_t* fc = o.spec_cnt + f;
        uint32_t emit_mask = 0u;
        // the 3*CPA class chunks alternate between the two warps that share a TMEM lane quarter (half 0 / 1); the next chunk's
        // TMEM read is in flight while the current one is compared
        int par = 0;
#pragma unroll 1
        for (int j = half; j < NCH; j += nhalf, par ^= 1) {
            const int a = j / CPA, cx = j - a * CPA;
            const int n = (REM != CH && cx == CPA - 1) ? REM : CH;
            const float la = (a == 0) ? ell[0] : ((a == 1) ? ell[1] : ell[2]);
            const float ca = (a == 0) ? conf[0] : ((a == 1) ? conf[1] : conf[2]);
            // first test on the RAW logits: x + b_i >= la needs x >= la - max b (a margin covers the roundings of both sides), so the
            // common case costs one compare per logit -- no add, no bias load
            const float bmax = scbias_s[NCH * CH4 + j];
            float thr = __fsub_rd(la, bmax);
            thr = __fsub_rd(thr, __fmaf_rn(1e-6f, __fadd_rn(fabsf(la), fabsf(bmax)), 1e-20f));
            tc::tmem_ld_wait();
            bool any = false;
            float xv[CH];
            if (par == 0) {
                if (j + nhalf < NCH) issue_j(j + nhalf, rc[1]);
#pragma unroll
                for (int i = 0; i < CH; ++i) { xv[i] = __uint_as_float(rc[0][i]); if (i < n) any |= xv[i] >= thr; }
            } else {
                if (j + nhalf < NCH) issue_j(j + nhalf, rc[0]);
#pragma unroll
                for (int i = 0; i < CH; ++i) { xv[i] = __uint_as_float(rc[1][i]); if (i < n) any |= xv[i] >= thr; }
            }
            if (any) {                                                    // rare: ~0.3 % of the class logits pass
                float bv[CH4];
#pragma unroll
                for (int i = 0; i < CH4; i += 4)
                    *reinterpret_cast<float4*>(bv + i) = *reinterpret_cast<const float4*>(scbias_s + j * CH4 + i);
#pragma unroll
                for (int i = 0; i < CH; ++i) bv[i] = __fadd_rn(xv[i], bv[i]);
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    if (i < n && bv[i] >= la && cx * CH + i < o.c_valid) {
                        const float scv = vd_score(bv[i], ca);
                        if (scv > vth) {
                            emit_mask |= 1u << a;
                            const uint32_t kh = __float_as_uint(scv) | 0x80000000u;
                            const uint32_t krow = row0 + (uint32_t)(cx * CH + i) * HW3 + (uint32_t)a;
                            VD_DEV_CHECK(krow < (uint32_t)o.rows_total && inb && f < o.frames);
                            const uint32_t pos = atomicAdd(fc, 1u);
                            if (pos < (uint32_t)kSpecCap) fl[pos] = ((uint64_t)kh << 32) | (uint32_t)~krow;
                        }
                    }
                }
            }
        }
        // raw box records of the emitting (pixel, anchor) pairs (decoded by the NMS kernel for the <= topk survivors)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const bool mine = ((emit_mask >> a) & 1u) != 0u;
            if (__any_sync(0xffffffffu, mine)) {
                uint32_t r4[4];
                tc::tmem_ld<4>(tb + (uint32_t)(a * P), r4); tc::tmem_ld_wait();
                if (mine) o.boxes[(size_t)f * o.anc_total + anc_base_s + cell * 3 + a] =
                    make_float4(__uint_as_float(r4[0]) + sbias_s[a * P + 0], __uint_as_float(r4[1]) + sbias_s[a * P + 1],
                                __uint_as_float(r4[2]) + sbias_s[a * P + 2], __uint_as_float(r4[3]) + sbias_s[a * P + 3]);
            }
        }
}

template <int C, int NPAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F_THREADS, 1)      // 168 registers: the register file is allocated as if for 12 warps (200 does not launch)
temporal_head_fused_kernel(const __grid_constant__ FusedMaps maps, const __grid_constant__ FusedParams p) {
    using Cfg = FusedCfg<C, NPAD>;
    constexpr int P = 5 + C;
    constexpr int KB4 = F_NT / F_BLOCK_K;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;
    unsigned char* stg = ring + Cfg::STAGES * Cfg::STAGE_BYTES;            // staged tip chunk: KB4 tiles [128 rows x 64 ch], swizzled
    float* sbn = reinterpret_cast<float*>(stg + Cfg::STG_BYTES);           // [scale][scale values (1024) | shift values (1024)]
    float* scbias = sbn + Cfg::BN_BYTES / 4;                               // [scale][3 anchors][CPA][CH4] class biases
    float* sbias = scbias + Cfg::CBIAS_BYTES / 4;                          // [scale][NPAD]
    FusedShared* sh = reinterpret_cast<FusedShared*>(sbias + VD_MAX_SCALES * NPAD);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;

    for (int s_ = 0; s_ < p.num_scales; ++s_)
        for (int i = threadIdx.x; i < p.sc[s_].Cin; i += F_THREADS) { sbn[s_ * 2048 + i] = p.sc[s_].scale[i]; sbn[s_ * 2048 + 1024 + i] = p.sc[s_].shift[i]; }
    for (int i = threadIdx.x; i < VD_MAX_SCALES * NPAD; i += F_THREADS) {
        const int s_ = i / NPAD, n = i % NPAD;
        sbias[i] = (s_ < p.num_scales && p.sc[s_].bias && n < 3 * P) ? p.sc[s_].bias[n] : 0.0f;
    }
    for (int s_ = 0; s_ < VD_MAX_SCALES; ++s_) spec_stage_tables<C>(scbias, s_ < p.num_scales ? p.sc[s_].bias : nullptr, s_, (int)threadIdx.x, F_THREADS);
    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) { tc::mbar_init(&sh->full[i], 1); tc::mbar_init(&sh->empty[i], 1); }
        tc::mbar_init(&sh->tip_full, 1); tc::mbar_init(&sh->tip_empty, 16);
        tc::mbar_init(&sh->stg_full, 16); tc::mbar_init(&sh->stg_empty, 1);
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&sh->pred_full[i], 1); tc::mbar_init(&sh->pred_empty[i], 16); }
        tc::fence_barrier_init();
        for (int s_ = 0; s_ < p.num_scales; ++s_) { tc::prefetch_tmap(&maps.x[s_]); tc::prefetch_tmap(&maps.w[s_]); tc::prefetch_tmap(&maps.wp[s_]); }
    }
    if (warp == 1) tc::tmem_alloc_2cta<512>(&sh->tmem_base);
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;
    // item of scale q -> (window b, row block mt) of CTA r; past the last tile: b == B (pure padding: zero rows in, nothing emitted)
    auto coords = [&](const FusedScale& q, int item, uint32_t r, int& b, int& mt) {
        const int m = item * 2 + (int)r;
        mt = m % q.m_tiles; b = m / q.m_tiles;
    };
    auto tap_active1 = [&](const FusedScale& q, int b, int mt, int dt) -> bool {
        const int r0 = mt * F_BLOCK_M;
        int r1 = r0 + F_BLOCK_M - 1; if (r1 > q.rows - 1) r1 = q.rows - 1;
        return (b < p.B) && (r1 + dt * q.HW >= 0) && (r0 + dt * q.HW <= q.rows - 1);
    };
    // k-th item of this CTA pair in scale s (see FusedParams::strided)
    auto n_items = [&](int s) -> int { return p.strided[s] + (int)p.beg[s][cluster_id + 1] - (int)p.beg[s][cluster_id]; };
    auto item_at = [&](int s, int k) -> int { return k < p.strided[s] ? k * p.pairs + cluster_id : (int)p.beg[s][cluster_id] + (k - p.strided[s]); };
    auto tap_active = [&](const FusedScale& q, int item, int dt) -> bool {
        int b0, m0, b1, m1; coords(q, item, 0, b0, m0); coords(q, item, 1, b1, m1);
        return tap_active1(q, b0, m0, dt) || tap_active1(q, b1, m1, dt);
    };

    if (warp == 0) {
        // =========================== TMA producer (both CTAs) ===========================
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            // Ring entries in the order the MMA role consumes them: the k-blocks of a chunk (A tile + half of the tap's weight tile), and,
            // behind the last k-block of the NEXT chunk, the prediction weights of the chunk before (KB4 tiles [WP_ROWS x 64 ch] in one
            // stage) -- that is where the MMA role issues that chunk's prediction GEMM.  No separate buffer, no extra barriers.
            bool pending = false; int p_s = 0, p_nt = 0;
            auto load_wp = [&]() {
                tc::mbar_wait_cluster(&sh->empty[stage], phase ^ 1u);
                unsigned char* dst = ring + stage * Cfg::STAGE_BYTES;
                if (rank == 0) tc::mbar_expect_tx(&sh->full[stage], 2u * Cfg::WP_BYTES);
                const uint32_t bar = tc::mapa_u32(&sh->full[stage], 0u);
#pragma unroll
                for (int j = 0; j < KB4; ++j)
                    tc::tma_load_2d_pair(dst + j * Cfg::WP_TILE, &maps.wp[p_s], bar, p_nt * F_NT + j * F_BLOCK_K, (int)rank * Cfg::WP_ROWS);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                pending = false;
            };
            for (int s = 0; s < p.num_scales; ++s) {
                const FusedScale& q = p.sc[s];
                const int kb_per_tap = q.Cin / F_BLOCK_K;
                for (int ik = 0, ni = n_items(s); ik < ni; ++ik) {
                    const int item = item_at(s, ik);
                    int b, mt; coords(q, item, rank, b, mt);
                    for (int nt = 0; nt < q.n_chunks; ++nt) {
                        for (int tap = 0; tap < 3; ++tap) {
                            const int dt = tap - 1;
                            if (!tap_active(q, item, dt)) continue;
                            for (int kb = 0; kb < kb_per_tap; ++kb) {
                                tc::mbar_wait_cluster(&sh->empty[stage], phase ^ 1u);
                                unsigned char* a_dst = ring + stage * Cfg::STAGE_BYTES;
                                if (rank == 0) tc::mbar_expect_tx(&sh->full[stage], 2u * Cfg::STAGE_BYTES);
                                const uint32_t bar = tc::mapa_u32(&sh->full[stage], 0u);
                                tc::tma_load_3d_pair(a_dst, &maps.x[s], bar, kb * F_BLOCK_K, mt * F_BLOCK_M + dt * q.HW, b);
                                tc::tma_load_3d_pair(a_dst + Cfg::A_BYTES, &maps.w[s], bar, kb * F_BLOCK_K, nt * F_NT + (int)rank * (F_NT / 2), tap);
                                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                            }
                        }
                        if (pending) load_wp();
                        pending = true; p_s = s; p_nt = nt;
                    }
                }
            }
            if (pending) load_wp();
        }
    } else if (warp == 1) {
        // =========================== MMA issuer (leader CTA) ===========================
        if (rank == 0 && tc::elect_one()) {
            constexpr uint32_t idesc_tip = tc::make_idesc_bf16(2 * F_BLOCK_M, F_NT);
            constexpr uint32_t idesc_pred = tc::make_idesc_bf16(2 * F_BLOCK_M, NPAD);
            int stage = 0; uint32_t phase = 0; uint32_t cc = 0, ic = 0;
            const uint32_t stg_addr = tc::smem_u32(stg);
            // The prediction GEMM of chunk c (A = the staged tip chunk of both CTAs, B = the chunk's prediction weights, which arrive
            // as a ring entry of their own) is issued right BEHIND the tip GEMM of chunk c + 1: the tip accumulator is single-buffered,
            // so while the epilogue warps read chunk c + 1 back (~1.3 k cycles) the tensor pipe would idle -- it runs these ~0.9 k
            // cycles of MMAs instead.  Chunk c was staged a whole chunk ago; its decode follows one chunk late as well (epilogue role).
            bool pending = false, p_last = false; uint32_t p_cc = 0, p_ic = 0; int p_nt = 0;
            auto issue_pred = [&]() {
                const uint32_t pb = p_ic & 1u;
                const uint32_t d_pred = tmem_base + 256u + pb * 128u;
                const bool st = p.stamps && blockIdx.x == 0 && p_cc < 200u;
                if (st) p.stamps[p_cc * 16 + 3] = clock64();
                tc::mbar_wait_cluster(&sh->full[stage], phase);                   // the chunk's prediction weights (ring entry)
                if (st) p.stamps[p_cc * 16 + 4] = clock64();
                tc::mbar_wait_cluster(&sh->stg_full, p_cc & 1u);
                if (p_nt == 0) tc::mbar_wait_cluster(&sh->pred_empty[pb], ((p_ic >> 1) & 1u) ^ 1u);
                tc::fence_after_sync();
                if (st) p.stamps[p_cc * 16 + 5] = clock64();
                const uint32_t wp_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
#pragma unroll
                for (int j = 0; j < KB4; ++j) {
                    const uint64_t da = tc::make_smem_desc_sw128(stg_addr + (uint32_t)(j * Cfg::STG_TILE));
                    const uint64_t db = tc::make_smem_desc_sw128(wp_addr + (uint32_t)(j * Cfg::WP_TILE));
#pragma unroll
                    for (int k = 0; k < F_BLOCK_K / 16; ++k)
                        tc::umma_bf16_2cta(d_pred, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_pred, (uint32_t)((p_nt | j | k) != 0));
                }
                tc::umma_commit_2cta(&sh->empty[stage]);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                tc::umma_commit_2cta(&sh->stg_empty);
                if (p_last) tc::umma_commit_2cta(&sh->pred_full[pb]);
                pending = false;
            };
            for (int s = 0; s < p.num_scales; ++s) {
                const FusedScale& q = p.sc[s];
                const int kb_per_tap = q.Cin / F_BLOCK_K;
                for (int ik = 0, ni = n_items(s); ik < ni; ++ik, ++ic) {
                    const int item = item_at(s, ik);
                    for (int nt = 0; nt < q.n_chunks; ++nt, ++cc) {
                        const bool st = p.stamps && blockIdx.x == 0 && cc < 200u;
                        if (st) p.stamps[cc * 16 + 0] = clock64();
                        tc::mbar_wait_cluster(&sh->tip_empty, (cc & 1u) ^ 1u);          // the epilogue has read the previous chunk's accumulator
                        tc::fence_after_sync();
                        if (st) p.stamps[cc * 16 + 1] = clock64();
                        uint32_t first = 1;
                        for (int tap = 0; tap < 3; ++tap) {
                            if (!tap_active(q, item, tap - 1)) continue;
                            for (int kb = 0; kb < kb_per_tap; ++kb) {
                                tc::mbar_wait_cluster(&sh->full[stage], phase);
                                tc::fence_after_sync();
                                const uint32_t a_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
                                const uint64_t da = tc::make_smem_desc_sw128(a_addr);
                                const uint64_t db = tc::make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
                                for (int k = 0; k < F_BLOCK_K / 16; ++k) {
                                    tc::umma_bf16_2cta(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_tip, first ? 0u : 1u);
                                    first = 0;
                                }
                                tc::umma_commit_2cta(&sh->empty[stage]);
                                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                            }
                        }
                        tc::umma_commit_2cta(&sh->tip_full);
                        if (st) p.stamps[cc * 16 + 2] = clock64();
                        if (pending) issue_pred();                 // the previous chunk's prediction GEMM runs while this accumulator is read back
                        pending = true; p_cc = cc; p_ic = ic; p_nt = nt; p_last = (nt == q.n_chunks - 1);
                    }
                }
            }
            if (pending) issue_pred();
        }
    } else {
        // =========================== epilogue warps (both CTAs) ===========================
        const int lq = warp & 3;                              // TMEM lane quarter
        const int half = (warp - 2) >> 2;                     // column half of the tip chunk / share of the class chunks
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        const int trow = lq * 32 + lane;                      // row inside the tile
        const uint32_t sw = (uint32_t)(trow & 7);
        const uint32_t stg_row = tc::smem_u32(stg) + (uint32_t)trow * 128u;
        const bool ws_ok = p.tile_counter[2] == p.ws_magic;   // the workspace holds thresholds of this layout
        SpecOut sout;
        sout.rows_total = p.g.row_base[p.g.num_scales]; sout.anc_total = p.g.anc_base[p.g.num_scales]; sout.c_valid = p.c_valid; sout.valid_thresh = p.valid_thresh; sout.boxes = p.boxes; sout.spec_lists = p.spec_lists;
        sout.spec_cnt = p.spec_cnt; sout.spec_tau = p.spec_tau; sout.frames = p.frames;
        // ---- decode + speculative candidate filter on an item's prediction accumulator (head_kernel<EPI_SPEC>, per-lane frame).  Runs one
        // chunk LATE: the item's last prediction GEMM is issued behind the NEXT chunk's tip GEMM (MMA role), i.e. while these warps read
        // that chunk's accumulator back and stage it; the decode of the previous item follows.
        auto decode_item = [&](const int s, const int item, const uint32_t ic, const uint32_t scc) {
            const FusedScale& q = p.sc[s];
            const int HW = q.HW;
            const float* sbias_s = sbias + s * NPAD;
            const float* scbias_s = scbias + s * SpecTables<C>::BLK;
            int b, mt; coords(q, item, rank, b, mt);
            const int row = mt * F_BLOCK_M + trow;
            const bool inb = (b < p.B) && (row < q.rows);
            const uint32_t pb = ic & 1u;
            const bool st2 = p.stamps && blockIdx.x == 0 && warp == 2 && lane == 0 && scc < 200u;
            const uint32_t tau_hint = (ws_ok && inb) ? __ldcg(p.spec_tau + (b * p.T + row / HW)) : 0u;
            tc::mbar_wait_cluster(&sh->pred_full[pb], (ic >> 1) & 1u);
            tc::fence_after_sync();
            if (st2) p.stamps[scc * 16 + 13] = clock64();
            if (!(p.dbg & 1)) {
                const int f = inb ? b * p.T + row / HW : 0;
                const int cell = inb ? row % HW : 0;
                spec_decode_lane<C>(tmem_base + 256u + pb * 128u + lane_addr, half, 2, inb, f, cell, p.g.row_base[s], p.g.anc_base[s], HW, sbias_s, scbias_s, ws_ok, tau_hint, sout);
            }
            if (st2) p.stamps[scc * 16 + 14] = clock64();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(&sh->pred_empty[pb], 0u);
        };
        uint32_t cc = 0, ic = 0;
        bool dec_pending = false; int d_s = 0, d_item = 0; uint32_t d_ic = 0, d_cc = 0;
        for (int s = 0; s < p.num_scales; ++s) {
        const FusedScale& q = p.sc[s];
        for (int ik = 0, ni = n_items(s); ik < ni; ++ik, ++ic) {
            const int item = item_at(s, ik);
            for (int nt = 0; nt < q.n_chunks; ++nt, ++cc) {
                const bool st = p.stamps && blockIdx.x == 0 && warp == 2 && lane == 0 && cc < 200u;
                tc::mbar_wait_cluster(&sh->tip_full, cc & 1u);
                tc::fence_after_sync();
                if (st) p.stamps[cc * 16 + 10] = clock64();
                // the warp's 128 accumulator columns go to registers in one go and the accumulator is handed back at once: it is
                // single-buffered, the tensor pipe waits for exactly this
                uint32_t r[128];
                const uint32_t tbase = tmem_base + (uint32_t)(half * 128) + lane_addr;
#pragma unroll
                for (int g = 0; g < 8; ++g) tc::tmem_ld16(tbase + (uint32_t)(g * 16), r + g * 16);
                tc::tmem_ld_wait();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive_cluster(&sh->tip_empty, 0u);
                tc::mbar_wait_cluster(&sh->stg_empty, (cc & 1u) ^ 1u);             // the previous chunk's prediction MMAs have read the staging tiles
                if (st) p.stamps[cc * 16 + 11] = clock64();
                const float2 slope2 = make_float2(p.slope, p.slope);
                const float4* sc = reinterpret_cast<const float4*>(sbn + s * 2048 + nt * F_NT + half * 128);
                const float4* sf = reinterpret_cast<const float4*>(sbn + s * 2048 + 1024 + nt * F_NT + half * 128);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const uint32_t dst = stg_row + (uint32_t)((half * 2 + t) * Cfg::STG_TILE);      // k-block (half*2+t) of the chunk
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint32_t pk[4];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = t * 64 + j * 8 + h * 4;
                            const float4 s4 = sc[i >> 2];
                            const float4 f4 = sf[i >> 2];
                            // packed fp32x2 FMA / MUL (sm_100: two lanes per instruction, each rounded like the scalar form); LeakyReLU as
                            // max(v, v * slope) (0 < slope < 1: the same bits as v > 0 ? v : v * slope)
                            const float2 v01 = __ffma2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), make_float2(s4.x, s4.y), make_float2(f4.x, f4.y));
                            const float2 v23 = __ffma2_rn(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), make_float2(s4.z, s4.w), make_float2(f4.z, f4.w));
                            const float2 m01 = __fmul2_rn(v01, slope2), m23 = __fmul2_rn(v23, slope2);
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(v01.x, m01.x), fmaxf(v01.y, m01.y)), h1 = __floats2bfloat162_rn(fmaxf(v23.x, m23.x), fmaxf(v23.y, m23.y));
                            pk[2 * h] = *reinterpret_cast<uint32_t*>(&h0); pk[2 * h + 1] = *reinterpret_cast<uint32_t*>(&h1);
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((uint32_t)j ^ sw) << 4)),
                                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    }
                }
                tc::fence_proxy_async_smem();                                     // st.shared -> visible to the UMMA (async proxy) reads
                __syncwarp();
                if (lane == 0) tc::mbar_arrive_cluster(&sh->stg_full, 0u);
                if (st) p.stamps[cc * 16 + 12] = clock64();
                if (dec_pending) { decode_item(d_s, d_item, d_ic, d_cc); dec_pending = false; }
            }
            dec_pending = true; d_s = s; d_item = item; d_ic = ic; d_cc = cc - 1u;
        }
        }
        if (dec_pending) decode_item(d_s, d_item, d_ic, d_cc);
    }
    __syncwarp();
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    if (warp == 1) tc::tmem_dealloc_2cta<512>(tmem_base);
}

// `items` equal items of `cost` each onto the current loads of `clusters` workers, each item to the least-loaded worker (longest-processing-
// time greedy when called for the costliest items first); cnt[c] = items worker c gets.  Levels are raised in bulk (a linear scan per
// item would be 4 k x 74 steps per call).
static void lpt_fill(double* load, int clusters, int items, double cost, int* cnt) {
    for (int c = 0; c < clusters; ++c) cnt[c] = 0;
    int left = items;
    while (left > 0) {
        int lo = 0;
        for (int c = 1; c < clusters; ++c) if (load[c] < load[lo]) lo = c;
        double next = 1e300;                     // the next higher load level
        int at_lo = 0;
        for (int c = 0; c < clusters; ++c) { if (load[c] <= load[lo] + 1e-9) ++at_lo; else if (load[c] < next) next = load[c]; }
        // every worker at the lowest level takes k items, k = what lifts it to the next level (at least 1), bounded by what is left
        long long k = next > 1e299 ? (left + at_lo - 1) / at_lo : (long long)((next - load[lo]) / cost);
        if (k < 1) k = 1;
        if (k * at_lo > left) k = left / at_lo;
        if (k < 1) {                              // fewer items left than workers at the level: one each
            for (int c = 0; c < clusters && left > 0; ++c) if (load[c] <= load[lo] + 1e-9) { ++cnt[c]; load[c] += cost; --left; }
            continue;
        }
        const double lvl = load[lo];
        for (int c = 0; c < clusters; ++c) if (load[c] <= lvl + 1e-9) { cnt[c] += (int)k; load[c] += k * cost; left -= (int)k; }
    }
}

// Deals the items of every scale to `clusters` CTA pairs: scales in the order given (s32, s16, s8 = decreasing item cost), each item
// to the pair with the least work so far; a pair's items of one scale are `strided` rounds + a contiguous range.
// Cost model of an item (k-cycles, from the per-chunk stamps of scripts/tfused_stamps.py): chunks x (0.7 per k-block + 1.5).
static void tfused_schedule(FusedParams* p, int clusters) {
    double load[F_MAX_CLUSTERS];
    int cnt[F_MAX_CLUSTERS];
    for (int c = 0; c < clusters; ++c) load[c] = 0.0;
    p->pairs = clusters;
    for (int s = 0; s < VD_MAX_SCALES; ++s) {
        for (int c = 0; c <= F_MAX_CLUSTERS; ++c) p->beg[s][c] = 0;
        p->strided[s] = 0;
        if (s >= p->num_scales) continue;
        const FusedScale& q = p->sc[s];
        const int items = q.m_tiles > 0 ? (int)(((long long)p->B * q.m_tiles + 1) / 2) : 0;
        const double cost = q.n_chunks * (0.7 * 3.0 * (q.Cin / F_BLOCK_K) + 1.5);
        lpt_fill(load, clusters, items, cost, cnt);
        int sr = 0;                                   // strided rounds: every pair takes part in them
        if (q.HW >= 8 * F_BLOCK_M) { sr = cnt[0]; for (int c = 1; c < clusters; ++c) if (cnt[c] < sr) sr = cnt[c]; }
        p->strided[s] = sr;
        int acc = sr * clusters;
        for (int c = 0; c < clusters; ++c) { p->beg[s][c] = (unsigned short)acc; acc += cnt[c] - sr; }
        for (int c = clusters; c <= F_MAX_CLUSTERS; ++c) p->beg[s][c] = (unsigned short)acc;
    }
}

template <int C, int NPAD>
static int launch_tfused_t(const FusedMaps& maps, const FusedParams& p, int clusters, cudaStream_t stream) {
    using Cfg = FusedCfg<C, NPAD>;
    auto kern = temporal_head_fused_kernel<C, NPAD>;
    { int rc_ = configure_kernel((const void*)kern, Cfg::SMEM_BYTES, false); if (rc_) return rc_; }
    if (clusters < 1) return VD_OK;
    kern<<<(unsigned)(2 * clusters), F_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, p);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

// class counts the fused kernel is compiled for (the VID / VOC heads); anything else runs the unfused chain
static bool tfused_supported(int C) { return C == 30 || C == 20; }

static int launch_tfused(const FusedMaps& maps, const FusedParams& p, int C, int clusters, cudaStream_t stream) {
    switch (C) {
        case 30: return launch_tfused_t<30, 112>(maps, p, clusters, stream);
        case 20: return launch_tfused_t<20, 80>(maps, p, clusters, stream);
        default: break;
    }
    return set_error(VD_ERR_UNSUPPORTED, "fused temporal head: no kernel shape for %d classes", C);
}

}  // namespace vd
