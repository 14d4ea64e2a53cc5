// head_pair_kernel -- the fused head (1x1 prediction conv -> YOLOOutputV3 decode -> speculative candidate filter; yolo3.py:62,157-199,523)
// on CTA PAIRS, for WIDE heads (256 prediction columns: COCO, num_class 31..80 on that shape).  Included by head.cu after tfused.cuh.
//
// Why: head_kernel's 1-CTA tiles pull the whole [256 x 64] weight tile + a [128 x 64] activation tile per k-block = 48 KB; at COCO-608 that
// is ~1.0 GB per 64-frame step, ~10 TB/s of L2 -> SM operand traffic -- the rate at which the tip-cell GEMM saturated --
// with the tensor pipe 49 % busy.  Here (tcgen05 cta_group::2, M256 x N256 x K16) each CTA stages its own 128 rows of activations and
// only HALF of the weight tile: 32 KB per k-block for the same MMA work, and a 6-stage ring instead of 3.
//   warp 0 (both CTAs)   TMA producer: A tile [128 rows x 64 ch] of the flattened (frames*HW) row axis + its half of the weights
//   warp 1 (leader CTA)  MMA issuer, two TMEM accumulators of 256 columns (double-buffered)
//   warps 2-17           decode + candidate filter of the CTA's 128 rows (spec_decode_lane of tfused.cuh: frame / cell per lane, the four
//                        warps of a TMEM lane quarter split the 15 class chunks)
// Items (two consecutive 128-row tiles of one scale) are dealt to the pairs on the host, largest first (lpt_ranges).  Candidate lists and
// box records are the ones head_kernel<EPI_SPEC> writes (order aside), so nms_spec_kernel and the exact fallback are unchanged.
#pragma once

namespace vd {

// EG = epilogue warps per TMEM lane quarter: they split the class chunks.  4 for the 256-column heads (240 class logits per pixel at C = 80: with 2
// the decode, not the mainloop, bound the kernel: 100 us per 64 frames against 58 us of mainloop); 2 for narrow heads.
constexpr int hp_threads(int EG) { return 64 + EG * 128; }

struct PairScale {
    int HW, Cin, rows, m_tiles;              // pixels per frame, channels, rows = frames*HW, 128-row tiles
    const float* bias;                       // (3*(5+C)) or null
};
struct PairParams {
    int num_scales, frames, pairs;
    PairScale sc[VD_MAX_SCALES];
    unsigned short beg[VD_MAX_SCALES][F_MAX_CLUSTERS + 1];   // pair c runs items beg[s][c] .. beg[s][c + 1] of scale s
    HeadGeom g;
    int c_valid; float valid_thresh;
    float4* boxes; uint64_t* spec_lists; uint32_t* spec_cnt; const uint32_t* spec_tau;
    const unsigned int* tile_counter; unsigned int ws_magic;
    int dbg;                                 // profiling aid (VD_HEAD_PAIR_DBG): 1 = skip the decode / filter epilogue
};
struct PairMaps { CUtensorMap a[VD_MAX_SCALES], w[VD_MAX_SCALES]; };

template <int C, int NPAD, int EG> struct PairCfg {
    static constexpr int THREADS = hp_threads(EG);
    static constexpr int ACC_STRIDE = (NPAD + 31) / 32 * 32;              // TMEM columns between the two accumulators
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE <= 256 ? 256 : 512;
    static constexpr int SMEM_BUDGET = (NPAD > 128 ? 212 : 200) * 1024;   // narrow heads leave the co-scheduled NMS CTA its 22 KB
    static constexpr int A_BYTES = 128 * 64 * 2;
    static constexpr int B_BYTES = (NPAD / 2) * 64 * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int CPA = (C + 15) / 16;
    static constexpr int CH = (C + CPA - 1) / CPA;
    static constexpr int CH4 = (CH + 3) / 4 * 4;
    static constexpr int CBIAS_BYTES = VD_MAX_SCALES * SpecTables<C>::BLK * 4;
    static constexpr int BIAS_BYTES = VD_MAX_SCALES * NPAD * 4;
    static constexpr int SH_BYTES = 1024;
    static constexpr int STAGES_RAW = (SMEM_BUDGET - CBIAS_BYTES - BIAS_BYTES - SH_BYTES - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + CBIAS_BYTES + BIAS_BYTES + SH_BYTES + 1024;
    static_assert(NPAD % 16 == 0 && (NPAD / 2) % 8 == 0 && NPAD <= 256 && 3 * (5 + C) <= NPAD, "prediction width");
    static_assert(STAGE_BYTES % 1024 == 0 && STAGES >= 3 && SMEM_BYTES <= 227 * 1024, "shared memory");
};
struct PairShared {
    uint64_t full[8], empty[8], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

template <int C, int NPAD, int EG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(hp_threads(EG), 1)
head_pair_kernel(const __grid_constant__ PairMaps maps, const __grid_constant__ PairParams p) {
    using Cfg = PairCfg<C, NPAD, EG>;
    constexpr int HP_THREADS = Cfg::THREADS;
    constexpr int HP_EPI_GROUPS = EG;
    constexpr int P = 5 + C;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;
    float* scbias = reinterpret_cast<float*>(ring + Cfg::STAGES * Cfg::STAGE_BYTES);      // [scale][3 anchors][CPA][CH4]
    float* sbias = scbias + Cfg::CBIAS_BYTES / 4;                                         // [scale][NPAD]
    PairShared* sh = reinterpret_cast<PairShared*>(sbias + VD_MAX_SCALES * NPAD);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;

    for (int i = threadIdx.x; i < VD_MAX_SCALES * NPAD; i += HP_THREADS) {
        const int s_ = i / NPAD, n = i % NPAD;
        sbias[i] = (s_ < p.num_scales && p.sc[s_].bias && n < 3 * P) ? p.sc[s_].bias[n] : 0.0f;
    }
    for (int s_ = 0; s_ < VD_MAX_SCALES; ++s_) spec_stage_tables<C>(scbias, s_ < p.num_scales ? p.sc[s_].bias : nullptr, s_, (int)threadIdx.x, HP_THREADS);
    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) { tc::mbar_init(&sh->full[i], 1); tc::mbar_init(&sh->empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&sh->acc_full[i], 1); tc::mbar_init(&sh->acc_empty[i], 8 * HP_EPI_GROUPS); }
        tc::fence_barrier_init();
        for (int s_ = 0; s_ < p.num_scales; ++s_) { tc::prefetch_tmap(&maps.a[s_]); tc::prefetch_tmap(&maps.w[s_]); }
    }
    if (warp == 1) tc::tmem_alloc_2cta<Cfg::TMEM_COLS>(&sh->tmem_base);
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int s = 0; s < p.num_scales; ++s) {
                const int nkb = p.sc[s].Cin / 64;
                for (int item = (int)p.beg[s][cluster_id]; item < (int)p.beg[s][cluster_id + 1]; ++item) {
                    const int m = item * 2 + (int)rank;                       // past the last tile: rows beyond the tensor read as zeros
                    for (int kb = 0; kb < nkb; ++kb) {
                        tc::mbar_wait_cluster(&sh->empty[stage], phase ^ 1u);
                        unsigned char* a_dst = ring + stage * Cfg::STAGE_BYTES;
                        if (rank == 0) tc::mbar_expect_tx(&sh->full[stage], 2u * Cfg::STAGE_BYTES);
                        const uint32_t bar = tc::mapa_u32(&sh->full[stage], 0u);
                        tc::tma_load_2d_pair(a_dst, &maps.a[s], bar, kb * 64, m * 128);
                        tc::tma_load_2d_pair(a_dst + Cfg::A_BYTES, &maps.w[s], bar, kb * 64, (int)rank * (NPAD / 2));
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(256, NPAD);
            int stage = 0; uint32_t phase = 0, it = 0;
            for (int s = 0; s < p.num_scales; ++s) {
                const int nkb = p.sc[s].Cin / 64;
                for (int item = (int)p.beg[s][cluster_id]; item < (int)p.beg[s][cluster_id + 1]; ++item, ++it) {
                    const uint32_t buf = it & 1u;
                    tc::mbar_wait_cluster(&sh->acc_empty[buf], ((it >> 1) & 1u) ^ 1u);
                    tc::fence_after_sync();
                    const uint32_t d_tmem = tmem_base + buf * Cfg::ACC_STRIDE;
                    for (int kb = 0; kb < nkb; ++kb) {
                        tc::mbar_wait_cluster(&sh->full[stage], phase);
                        tc::fence_after_sync();
                        const uint32_t a_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
                        const uint64_t da = tc::make_smem_desc_sw128(a_addr);
                        const uint64_t db = tc::make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc::umma_bf16_2cta(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                        tc::umma_commit_2cta(&sh->empty[stage]);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                    tc::umma_commit_2cta(&sh->acc_full[buf]);
                }
            }
        }
    } else {
        const int lq = warp & 3;
        const int half = (warp - 2) >> 2;
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        const int trow = lq * 32 + lane;
        const bool ws_ok = p.tile_counter[2] == p.ws_magic;
        SpecOut sout;
        sout.rows_total = p.g.row_base[p.g.num_scales]; sout.anc_total = p.g.anc_base[p.g.num_scales]; sout.c_valid = p.c_valid;
        sout.valid_thresh = p.valid_thresh; sout.boxes = p.boxes; sout.spec_lists = p.spec_lists; sout.spec_cnt = p.spec_cnt;
        sout.spec_tau = p.spec_tau; sout.frames = p.frames;
        uint32_t it = 0;
        for (int s = 0; s < p.num_scales; ++s) {
            const int HW = p.sc[s].HW, rows = p.sc[s].rows;
            const float* sbias_s = sbias + s * NPAD;
            const float* scbias_s = scbias + s * SpecTables<C>::BLK;
            const int row_base_s = p.g.row_base[s], anc_base_s = p.g.anc_base[s];
            for (int item = (int)p.beg[s][cluster_id]; item < (int)p.beg[s][cluster_id + 1]; ++item, ++it) {
                const uint32_t buf = it & 1u;
                const int row = (item * 2 + (int)rank) * 128 + trow;
                const bool inb = row < rows;
                const int f = inb ? row / HW : 0, cell = inb ? row - (row / HW) * HW : 0;
                const uint32_t tau_hint = ws_ok ? __ldcg(p.spec_tau + f) : 0u;       // its L2 latency hides behind the wait for the accumulator
                tc::mbar_wait_cluster(&sh->acc_full[buf], (it >> 1) & 1u);
                tc::fence_after_sync();
                if (!(p.dbg & 1))
                    spec_decode_lane<C>(tmem_base + buf * Cfg::ACC_STRIDE + lane_addr, half, HP_EPI_GROUPS, inb, f, cell, row_base_s, anc_base_s, HW, sbias_s, scbias_s, ws_ok, tau_hint, sout);
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive_cluster(&sh->acc_empty[buf], 0u);
            }
        }
    }
    __syncwarp();
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    if (warp == 1) tc::tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
}

template <int C, int NPAD, int EG>
static int launch_hpair_t(const PairMaps& maps, const PairParams& p, int clusters, cudaStream_t stream) {
    using Cfg = PairCfg<C, NPAD, EG>;
    auto kern = head_pair_kernel<C, NPAD, EG>;
    { int rc_ = configure_kernel((const void*)kern, Cfg::SMEM_BYTES, NPAD <= 128); if (rc_) return rc_; }      // narrow heads: same carve-out as the NMS kernel that shares the SM
    if (clusters < 1) return VD_OK;
    kern<<<(unsigned)(2 * clusters), Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(maps, p);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

// Compiled for the 256-column shape only.  The narrow VOC / VID shapes (<20,80,2>, <30,112,2>, NMS CTA co-resident) were measured and
// rejected: 48.7 us per VOC step against 33.9 us for head_kernel (its dynamic tile scheduler, three epilogue groups with per-warp
// staging and the 7-stage ring are what an HBM-bound head needs; the operand traffic this kernel saves does not bind there).
static bool hpair_supported(int C) { return C == 80; }

static int launch_hpair(const PairMaps& maps, const PairParams& p, int C, int clusters, cudaStream_t stream) {
    if (C == 80) return launch_hpair_t<80, 256, 4>(maps, p, clusters, stream);
    return set_error(VD_ERR_UNSUPPORTED, "pair head: no kernel shape for %d classes", C);
}

}  // namespace vd
