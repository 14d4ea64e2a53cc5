// vd_temporal_conv -- the temporal (3,1,1) cell of the (2+1)D tip:
//   Conv3D(C -> C, kernel (3,1,1), pad (1,0,0), stride 1, no bias) + BatchNorm(eps 1e-5) + LeakyReLU(0.1)
//   (models/definitions/layers.py:82-89 `_conv21d` second cell -> :73-79 `_conv3d`; used as the
//   detection block's tip at yolo3_temporal.py:226-227, whose output feeds YOLOOutputV3 directly).
//
// Implicit GEMM on tcgen05:  Y[b, r, :] = lrelu(scale * sum_{dt=-1..1} X[b, r + dt*HW, :] @ W_dt^T + shift)
// with r = t*HW + pixel flattened over the window, so the zero padding in time IS the TMA
// out-of-bounds fill of the 3-D map (C, T*HW, B): rows r + dt*HW outside [0, T*HW) read as zeros,
// and taps that are out of range for a whole tile are skipped (13 of 15 taps do work at T=5).
//   warp 0: TMA producer (A tile [128 rows x 64 ch] of the shifted frame + W_dt tile [NT x 64])
//   warp 1: MMA issuer, M=128 x N=NT x K=16, accumulators double-buffered in TMEM
//   warps 2-9: epilogue (two warpgroups, each takes half of the tile's columns): tcgen05.ld -> folded BN (scale/shift
//              staged in shared memory) -> LeakyReLU -> bf16 -> 128-bit stores
#include <stdlib.h>
#include "tc.cuh"

namespace vd {

constexpr int T_BLOCK_M = 128;
constexpr int T_BLOCK_K = 64;
constexpr int T_THREADS = 320;               // TMA warp, MMA warp, 8 epilogue warps

struct TConvParams {
    int B, T, HW, C, rows;         // rows = T*HW
    int m_tiles, n_tiles, total_tiles;
    const float* scale; const float* shift; float slope;
    __nv_bfloat16* y;
    // fp32-parity modes (vd_temporal_conv_ex): x (planes, B, T*HW, C), w (planes, 3, Cout, Cin), y (planes, B, T*HW, C); the
    // K loop of a tap runs n_prod plane products (a_pl[i], w_pl[i]); planes == 1: the plain bf16 cell
    int planes, n_prod;
    signed char a_pl[8], w_pl[8];
    long long y_plane_stride;      // elements between output planes
    int prefetch;                  // pair kernel: L2 prefetch of the next item's activation boxes (VD_TCONV_PREFETCH, default on)
    const unsigned int* frame_list; // pair kernel, with cond: only the windows of the listed frames (frame_list[0 .. *cond)) are computed -- the exact
                                   // fallback of the fused temporal head needs the tip of the frames it redoes, nothing else
    const unsigned int* cond;      // non-null: the launch does nothing when *cond == 0 (the fused head's exact fallback recomputes the tip only when frames failed)
    int dbg;                       // profiling aid (VD_TCONV_DBG=1): the epilogue releases its accumulator without BN / stores, results are garbage
};
struct TConvMaps { CUtensorMap x; CUtensorMap w; CUtensorMap y; };      // y: output map of the pair kernel's TMA-store epilogue, box {64 ch, 128 rows, 1}

template <int NT> struct TConvCfg {
    static constexpr int A_BYTES = T_BLOCK_M * T_BLOCK_K * 2;
    static constexpr int B_BYTES = NT * T_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    static constexpr int BN_BYTES = 2 * 1024 * 4;        // folded BN scale / shift of up to 1024 channels
    static constexpr int TMEM_COLS = 2 * NT;            // NT in {128, 256}
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BN_BYTES + 2048 + 1024;
};

struct TShared {
    uint64_t full[8], empty[8], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

template <int NT>
__global__ void __launch_bounds__(T_THREADS, 1)
temporal_conv_kernel(const __grid_constant__ TConvMaps maps, const __grid_constant__ TConvParams p) {
    using Cfg = TConvCfg<NT>;
    if (p.cond && *p.cond == 0u) return;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;
    float* sscale = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);      // [C] then shift [C]
    float* sshift = sscale + 1024;
    TShared* sh = reinterpret_cast<TShared*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BN_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < p.C; i += T_THREADS) { sscale[i] = p.scale[i]; sshift[i] = p.shift[i]; }

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) { tc::mbar_init(&sh->full[i], 1); tc::mbar_init(&sh->empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&sh->tmem_full[i], 1); tc::mbar_init(&sh->tmem_empty[i], 8); }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&maps.x); tc::prefetch_tmap(&maps.w);
    }
    if (warp == 1) tc::tmem_alloc<Cfg::TMEM_COLS>(&sh->tmem_base);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;
    const int kb_per_tap = p.C / T_BLOCK_K;

    // tile -> (window b, row block mt, channel block nt); nt fastest so A tiles are re-read from L2
    auto coords = [&](int tile, int& b, int& mt, int& nt) {
        nt = tile % p.n_tiles; int r = tile / p.n_tiles; mt = r % p.m_tiles; b = r / p.m_tiles;
    };
    auto tap_active = [&](int mt, int dt) -> bool {   // does any row of the tile see frame t+dt inside the window?
        const int r0 = mt * T_BLOCK_M;
        int r1 = r0 + T_BLOCK_M - 1; if (r1 > p.rows - 1) r1 = p.rows - 1;
        return (r1 + dt * p.HW >= 0) && (r0 + dt * p.HW <= p.rows - 1);
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int b, mt, nt; coords(tile, b, mt, nt);
                // item order (shared with the MMA role): plain cell: tap -> k-block.  Parity modes: first every low-order plane product
                // (tap -> product -> k-block), then the p0 w0 product (tap -> k-block), see the MMA role for the chunking.
                const int phases = p.planes > 1 ? 2 : 1;
                for (int ph = 0; ph < phases; ++ph) {
                    const int pr0 = (p.planes > 1 && ph == 1) ? p.n_prod - 1 : 0;
                    const int pr1 = (p.planes > 1 && ph == 0) ? p.n_prod - 1 : p.n_prod;
                    for (int tap = 0; tap < 3; ++tap) {
                        const int dt = tap - 1;
                        if (!tap_active(mt, dt)) continue;
                        for (int pr = pr0; pr < pr1; ++pr) {
                            const int xb = b + (int)p.a_pl[pr] * p.B, wt = tap + 3 * (int)p.w_pl[pr];     // plane-major operands
                            for (int kb = 0; kb < kb_per_tap; ++kb) {
                                tc::mbar_wait(&sh->empty[stage], phase ^ 1u);
                                unsigned char* a_dst = ring + stage * Cfg::STAGE_BYTES;
                                tc::mbar_expect_tx(&sh->full[stage], Cfg::STAGE_BYTES);
                                tc::tma_load_3d(a_dst, &maps.x, &sh->full[stage], kb * T_BLOCK_K, mt * T_BLOCK_M + dt * p.HW, xb);
                                tc::tma_load_3d(a_dst + Cfg::A_BYTES, &maps.w, &sh->full[stage], kb * T_BLOCK_K, nt * NT, wt);
                                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(T_BLOCK_M, NT);
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;      // `it` counts accumulator hand-overs (chunks): one per tile in the plain cell
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int b, mt, nt; coords(tile, b, mt, nt);
                int n_act = 0;
                for (int tap = 0; tap < 3; ++tap) n_act += tap_active(mt, tap - 1) ? 1 : 0;
                // chunks of k-blocks that accumulate into one TMEM buffer before the epilogue takes it.  Plain cell: the whole tile.
                // Parity modes: chunk 0 = every low-order plane product (their truncation errors scale with their 2^-9 / 2^-18
                // magnitude), then the p0 w0 product in chunks of 4 k-blocks = 16 accumulates; the epilogue adds the chunks in
                // registers with round-to-nearest (the tensor core's fp32 accumulator truncates: error grows linearly with the chain).
                int lows = p.planes > 1 ? n_act * (p.n_prod - 1) * kb_per_tap : 0;
                int his = n_act * kb_per_tap;
                while (lows > 0 || his > 0) {
                    int n_items;
                    if (lows > 0) { n_items = lows; lows = 0; }
                    else if (p.planes > 1) { n_items = his < 4 ? his : 4; his -= n_items; }
                    else { n_items = his; his = 0; }
                    const uint32_t buf = it & 1u;
                    tc::mbar_wait(&sh->tmem_empty[buf], ((it >> 1) & 1u) ^ 1u);
                    tc::fence_after_sync();
                    const uint32_t d_tmem = tmem_base + buf * NT;
                    uint32_t first = 1;
                    for (int kb = 0; kb < n_items; ++kb) {
                        tc::mbar_wait(&sh->full[stage], phase);
                        tc::fence_after_sync();
                        const uint32_t a_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
                        const uint64_t da = tc::make_smem_desc_sw128(a_addr);
                        const uint64_t db = tc::make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < T_BLOCK_K / 16; ++k) {
                            tc::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
                            first = 0;
                        }
                        tc::umma_commit(&sh->empty[stage]);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                    tc::umma_commit(&sh->tmem_full[buf]);
                    ++it;
                }
            }
        }
    } else {
        const int q = warp & 3;                               // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                     // which half of the tile's columns
        constexpr int NH = NT / 2;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t it = 0;
        if constexpr (NT == 128) {
            if (p.planes > 1) {
                // parity modes (always on 128-wide channel blocks): add the tile's chunks in registers, then BN + LeakyReLU and the plane split
                for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                    int b, mt, nt; coords(tile, b, mt, nt);
                    const int row = mt * T_BLOCK_M + q * 32 + lane;
                    const bool inb = row < p.rows;
                    int n_act = 0;
                    for (int tap = 0; tap < 3; ++tap) n_act += tap_active(mt, tap - 1) ? 1 : 0;
                    const int n_chunks = 1 + (n_act * kb_per_tap + 3) / 4;
                    float sum[NH];
#pragma unroll
                    for (int i = 0; i < NH; ++i) sum[i] = 0.0f;
                    for (int c = 0; c < n_chunks; ++c, ++it) {
                        const uint32_t buf = it & 1u;
                        tc::mbar_wait(&sh->tmem_full[buf], (it >> 1) & 1u);
                        tc::fence_after_sync();
                        const uint32_t tb = tmem_base + buf * NT + (uint32_t)(half * NH) + lane_addr;
#pragma unroll
                        for (int n0 = 0; n0 < NH; n0 += 32) {
                            uint32_t r[32];
                            tc::tmem_ld16(tb + n0, r); tc::tmem_ld16(tb + n0 + 16, r + 16); tc::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) sum[n0 + i] = __fadd_rn(sum[n0 + i], __uint_as_float(r[i]));
                        }
                        tc::fence_before_sync();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
                    }
                    const int col0 = nt * NT + half * NH;
                    __nv_bfloat16* yrow = p.y + ((size_t)b * p.rows + (inb ? row : 0)) * p.C + col0;
#pragma unroll
                    for (int i = 0; i < NH; ++i) {        // the reference's BatchNorm roundings: (y * scale) + shift, then LeakyReLU
                        float v = __fadd_rn(__fmul_rn(sum[i], sscale[col0 + i]), sshift[col0 + i]);
                        sum[i] = v > 0.f ? v : v * p.slope;
                    }
#pragma unroll 1
                    for (int pl = 0; pl < p.planes; ++pl) {
#pragma unroll
                        for (int n0 = 0; n0 < NH; n0 += 8) {
                            uint32_t packed[4];
#pragma unroll
                            for (int i = 0; i < 8; i += 2) {
                                __nv_bfloat162 h = __floats2bfloat162_rn(sum[n0 + i], sum[n0 + i + 1]);
                                packed[i / 2] = *reinterpret_cast<uint32_t*>(&h);
                                sum[n0 + i] = __fsub_rn(sum[n0 + i], __low2float(h));             // exact: what the next plane carries
                                sum[n0 + i + 1] = __fsub_rn(sum[n0 + i + 1], __high2float(h));
                            }
                            if (inb) *reinterpret_cast<uint4*>(yrow + (size_t)pl * p.y_plane_stride + n0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        }
                    }
                }
                it = 0xffffffffu;      // nothing left for the plain loop below
            }
        }
        for (int tile = blockIdx.x; tile < p.total_tiles && it != 0xffffffffu; tile += gridDim.x, ++it) {
            int b, mt, nt; coords(tile, b, mt, nt);
            const uint32_t buf = it & 1u;
            const int row = mt * T_BLOCK_M + q * 32 + lane;
            const bool inb = row < p.rows;
            tc::mbar_wait(&sh->tmem_full[buf], (it >> 1) & 1u);
            tc::fence_after_sync();
            const uint32_t tbase = tmem_base + buf * NT + (uint32_t)(half * NH) + lane_addr;
            const int col0 = nt * NT + half * NH;
            __nv_bfloat16* yrow = p.y + ((size_t)b * p.rows + (inb ? row : 0)) * p.C + col0;
            const float* sc = sscale + col0;
            const float* sf = sshift + col0;
#pragma unroll 1
            for (int n0 = 0; n0 < NH; n0 += 32) {
                uint32_t r[32];
                tc::tmem_ld16(tbase + n0, r); tc::tmem_ld16(tbase + n0 + 16, r + 16); tc::tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 s4 = *reinterpret_cast<const float4*>(sc + n0 + i);
                    const float4 f4 = *reinterpret_cast<const float4*>(sf + n0 + i);
                    float v0 = fmaf(__uint_as_float(r[i]), s4.x, f4.x), v1 = fmaf(__uint_as_float(r[i + 1]), s4.y, f4.y);
                    float v2 = fmaf(__uint_as_float(r[i + 2]), s4.z, f4.z), v3 = fmaf(__uint_as_float(r[i + 3]), s4.w, f4.w);
                    v0 = v0 > 0.f ? v0 : v0 * p.slope; v1 = v1 > 0.f ? v1 : v1 * p.slope;
                    v2 = v2 > 0.f ? v2 : v2 * p.slope; v3 = v3 > 0.f ? v3 : v3 * p.slope;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
                    packed[i / 2] = *reinterpret_cast<uint32_t*>(&h0); packed[i / 2 + 1] = *reinterpret_cast<uint32_t*>(&h1);
                }
                if (inb) {
                    uint4* dst = reinterpret_cast<uint4*>(yrow + n0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
        }
    }
    __syncwarp();                // warps 0 / 1 ran single-lane role loops: reconverge before the aligned CTA barrier
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant for 256-wide channel blocks (tcgen05 cta_group::2; same protocol as conv_bn_lrelu_pair_kernel in
// conv.cu): the two CTAs of a cluster take two consecutive 128-row tiles of the flattened (window, T*HW) row axis and the same
// 256 output channels; ONE M256 x N256 x K16 MMA stream is issued by the leader; each CTA stages its own A tile and HALF of
// the tap's weight tile, both CTAs' TMA loads report their bytes to the leader's full barrier, commits are multicast.
// ---------------------------------------------------------------------------------------------------------------
struct T2Shared {
    uint64_t full[8], empty[8], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};
#ifndef VD_TCONV_OUT_BUFS
#define VD_TCONV_OUT_BUFS 2
#endif
struct T2Cfg {
    static constexpr int NT = 256;
    static constexpr int A_BYTES = T_BLOCK_M * T_BLOCK_K * 2;
    static constexpr int B_BYTES = (NT / 2) * T_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // output staging of the TMA-store epilogue: per column half (4 epilogue warps) OUT_BUFS tiles of [128 rows x 64 ch] bf16 in the
    // 128-byte-swizzled layout of a TMA box
    static constexpr int OUT_BUFS = VD_TCONV_OUT_BUFS;
    static constexpr int OUT_TILE_BYTES = T_BLOCK_M * 64 * 2;
    static constexpr int OUT_BYTES = 2 * OUT_BUFS * OUT_TILE_BYTES;
    static constexpr int STAGES = (212 * 1024 - OUT_BYTES) / STAGE_BYTES > 8 ? 8 : (212 * 1024 - OUT_BYTES) / STAGE_BYTES;
    static constexpr int BN_BYTES = 2 * 1024 * 4;
    static constexpr int TMEM_COLS = 2 * NT;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + BN_BYTES + 2048 + 1024;
    static_assert(OUT_BUFS == 1 || OUT_BUFS == 2, "OUT_BUFS");
    static_assert(SMEM_BYTES <= 227 * 1024 && STAGES >= 3, "shared memory");
};
// named barrier of one column half's 4 epilogue warps (immediate ids, like head.cu)
__device__ __forceinline__ void t2_half_bar(int half) {
    if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T_THREADS, 1)
temporal_conv_pair_kernel(const __grid_constant__ TConvMaps maps, const __grid_constant__ TConvParams p) {
    using Cfg = T2Cfg;
    constexpr int NT = Cfg::NT;
    if (p.cond && *p.cond == 0u) return;                         // both CTAs of the pair read the same word
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;
    unsigned char* outb = smem + Cfg::STAGES * Cfg::STAGE_BYTES;                 // [half][OUT_BUFS][128 rows x 128 B], 1024-byte aligned
    float* sscale = reinterpret_cast<float*>(outb + Cfg::OUT_BYTES);
    float* sshift = sscale + 1024;
    T2Shared* sh = reinterpret_cast<T2Shared*>(outb + Cfg::OUT_BYTES + Cfg::BN_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    for (int i = threadIdx.x; i < p.C; i += T_THREADS) { sscale[i] = p.scale[i]; sshift[i] = p.shift[i]; }

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) { tc::mbar_init(&sh->full[i], 1); tc::mbar_init(&sh->empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&sh->tmem_full[i], 1); tc::mbar_init(&sh->tmem_empty[i], 16); }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&maps.x); tc::prefetch_tmap(&maps.w); tc::prefetch_tmap(&maps.y);
    }
    if (warp == 1) tc::tmem_alloc_2cta<Cfg::TMEM_COLS>(&sh->tmem_base);
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;
    const int kb_per_tap = p.C / T_BLOCK_K;
    const int m_total = p.B * p.m_tiles;                       // 128-row tiles over all windows
    const int pairs_per_win = (p.m_tiles + 1) >> 1;
    const int n_listed = p.frame_list ? (int)*p.cond : 0;      // list mode: windows of the listed frames only (pairs do not straddle windows)
    const int total_pairs = p.frame_list ? n_listed * pairs_per_win * p.n_tiles : ((m_total + 1) >> 1) * p.n_tiles;

    // pair -> (window b, row block mt) of CTA r, channel block nt; past the last tile: b == B (pure padding, nothing stored)
    auto coords = [&](int pair, uint32_t r, int& b, int& mt, int& nt) {
        nt = pair % p.n_tiles;
        if (p.frame_list) {
            const int qq = pair / p.n_tiles, wi = qq / pairs_per_win;
            mt = (qq - wi * pairs_per_win) * 2 + (int)r;
            b = mt < p.m_tiles ? (int)(p.frame_list[wi] / (unsigned)p.T) : p.B;
            if (mt >= p.m_tiles) mt = 0;
            return;
        }
        const int m = (pair / p.n_tiles) * 2 + (int)r;
        mt = m % p.m_tiles; b = m / p.m_tiles;
    };
    auto tap_active1 = [&](int b, int mt, int dt) -> bool {
        const int r0 = mt * T_BLOCK_M;
        int r1 = r0 + T_BLOCK_M - 1; if (r1 > p.rows - 1) r1 = p.rows - 1;
        return (b < p.B) && (r1 + dt * p.HW >= 0) && (r0 + dt * p.HW <= p.rows - 1);
    };
    auto tap_active = [&](int pair, int dt) -> bool {          // skipped only if pure padding for BOTH tiles of the pair
        int b0, m0, n0, b1, m1, n1; coords(pair, 0, b0, m0, n0); coords(pair, 1, b1, m1, n1);
        return tap_active1(b0, m0, dt) || tap_active1(b1, m1, dt);
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
                int b, mt, nt; coords(pair, rank, b, mt, nt);
                // L2 prefetch of the NEXT item's activation boxes, one per k-block of this item: the shared-memory ring (4-5 stages =
                // ~1 us of MMA work) covers an L2 hit but not a first touch that goes to HBM under load
                int b2 = p.B, mt2 = 0, nt2 = 0;
                if (p.prefetch && pair + num_clusters < total_pairs) coords(pair + num_clusters, rank, b2, mt2, nt2);
                for (int tap = 0; tap < 3; ++tap) {
                    const int dt = tap - 1;
                    if (!tap_active(pair, dt)) continue;
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        tc::mbar_wait_cluster(&sh->empty[stage], phase ^ 1u);
                        unsigned char* a_dst = ring + stage * Cfg::STAGE_BYTES;
                        if (rank == 0) tc::mbar_expect_tx(&sh->full[stage], 2u * Cfg::STAGE_BYTES);
                        const uint32_t bar = tc::mapa_u32(&sh->full[stage], 0u);
                        if (b2 < p.B) tc::tma_prefetch_3d(&maps.x, kb * T_BLOCK_K, mt2 * T_BLOCK_M + dt * p.HW, b2);
                        tc::tma_load_3d_pair(a_dst, &maps.x, bar, kb * T_BLOCK_K, mt * T_BLOCK_M + dt * p.HW, b);
                        tc::tma_load_3d_pair(a_dst + Cfg::A_BYTES, &maps.w, bar, kb * T_BLOCK_K, nt * NT + (int)rank * (NT / 2), tap);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(2 * T_BLOCK_M, NT);
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            for (int pair = cluster_id; pair < total_pairs; pair += num_clusters, ++it) {
                const uint32_t buf = it & 1u;
                tc::mbar_wait_cluster(&sh->tmem_empty[buf], ((it >> 1) & 1u) ^ 1u);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * NT;
                uint32_t first = 1;
                for (int tap = 0; tap < 3; ++tap) {
                    if (!tap_active(pair, tap - 1)) continue;
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        tc::mbar_wait_cluster(&sh->full[stage], phase);
                        tc::fence_after_sync();
                        const uint32_t a_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
                        const uint64_t da = tc::make_smem_desc_sw128(a_addr);
                        const uint64_t db = tc::make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < T_BLOCK_K / 16; ++k) {
                            tc::umma_bf16_2cta(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
                            first = 0;
                        }
                        tc::umma_commit_2cta(&sh->empty[stage]);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
                tc::umma_commit_2cta(&sh->tmem_full[buf]);
            }
        }
    } else {
        // Epilogue: 8 warps; warp (q, half) owns TMEM lanes q*32.. (rows) and 128 of the tile's 256 columns, in two passes of 64.
        // Per pass: tcgen05.ld -> folded BN -> LeakyReLU -> bf16 -> st.shared into the half's staging tile (swizzled like a TMA box:
        // conflict-free 128-bit stores) -> fence.proxy.async + named barrier of the half -> ONE thread issues a TMA store of the
        // [128 rows x 64 ch] box (rows past the window's end are clipped by the map).  The direct per-thread global stores this replaces
        // (32 rows x 16 B per instruction) kept the LSU busy for ~4 k cycles per tile: s8 ran at 0.31 ms against 0.20 ms of mainloop.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int NH = NT / 2;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const bool issuer = (q == 0 && lane == 0);
        const int trow = q * 32 + lane;                                       // row inside the tile
        unsigned char* obase = outb + half * Cfg::OUT_BUFS * Cfg::OUT_TILE_BYTES;
        const uint32_t orow = tc::smem_u32(obase) + (uint32_t)trow * 128u;
        const uint32_t sw = (uint32_t)(trow & 7);
        uint32_t it = 0, ob = 0;
        for (int pair = cluster_id; pair < total_pairs; pair += num_clusters, ++it) {
            int b, mt, nt; coords(pair, rank, b, mt, nt);
            const uint32_t buf = it & 1u;
            tc::mbar_wait_cluster(&sh->tmem_full[buf], (it >> 1) & 1u);
            tc::fence_after_sync();
            const uint32_t tbase = tmem_base + buf * NT + (uint32_t)(half * NH) + lane_addr;
            const int col0 = nt * NT + half * NH;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass, ++ob) {
                uint32_t r[64];
                const uint32_t ta = tbase + (uint32_t)(pass * 64);
                tc::tmem_ld16(ta, r); tc::tmem_ld16(ta + 16, r + 16); tc::tmem_ld16(ta + 32, r + 32); tc::tmem_ld16(ta + 48, r + 48);
                tc::tmem_ld_wait();
                if (pass == 1) {                                              // this warp has read its share of the accumulator
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive_cluster(&sh->tmem_empty[buf], 0u);
                }
                if (p.dbg & 1) continue;
                const float* sc = sscale + col0 + pass * 64;
                const float* sf = sshift + col0 + pass * 64;
                uint32_t packed[32];
#pragma unroll
                for (int i = 0; i < 64; i += 4) {
                    const float4 s4 = *reinterpret_cast<const float4*>(sc + i);
                    const float4 f4 = *reinterpret_cast<const float4*>(sf + i);
                    float v0 = fmaf(__uint_as_float(r[i]), s4.x, f4.x), v1 = fmaf(__uint_as_float(r[i + 1]), s4.y, f4.y);
                    float v2 = fmaf(__uint_as_float(r[i + 2]), s4.z, f4.z), v3 = fmaf(__uint_as_float(r[i + 3]), s4.w, f4.w);
                    v0 = v0 > 0.f ? v0 : v0 * p.slope; v1 = v1 > 0.f ? v1 : v1 * p.slope;
                    v2 = v2 > 0.f ? v2 : v2 * p.slope; v3 = v3 > 0.f ? v3 : v3 * p.slope;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
                    packed[i / 2] = *reinterpret_cast<uint32_t*>(&h0); packed[i / 2 + 1] = *reinterpret_cast<uint32_t*>(&h1);
                }
                const uint32_t slot = Cfg::OUT_BUFS == 2 ? (ob & 1u) : 0u;
                if (Cfg::OUT_BUFS == 1) {                                     // single staging tile: wait until the previous store has read it
                    if (issuer) tc::tma_store_wait_read<0>();
                    __syncwarp();
                    t2_half_bar(half);
                }
                const uint32_t dst = orow + slot * (uint32_t)Cfg::OUT_TILE_BYTES;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((uint32_t)j ^ sw) << 4)),
                                 "r"(packed[4 * j]), "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3]) : "memory");
                tc::fence_proxy_async_smem();
                // two staging tiles: the store issued one pass ago must have read its tile before anyone passes this barrier, because
                // the NEXT pass overwrites that tile
                if (Cfg::OUT_BUFS == 2 && issuer) tc::tma_store_wait_read<0>();
                __syncwarp();
                t2_half_bar(half);
                if (issuer && b < p.B) {
                    tc::tma_store_3d(&maps.y, obase + slot * Cfg::OUT_TILE_BYTES, col0 + pass * 64, mt * T_BLOCK_M, b);
                    tc::tma_store_commit();
                }
            }
        }
        if (issuer) tc::tma_store_wait<0>();
    }
    __syncwarp();
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    if (warp == 1) tc::tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
}

static int launch_tconv_pair(const TConvMaps& maps, const TConvParams& p, cudaStream_t stream) {
    auto kern = temporal_conv_pair_kernel;
    { int rc_ = configure_kernel((const void*)kern, T2Cfg::SMEM_BYTES, false); if (rc_) return rc_; }
    const long long pairs = ((long long)p.B * p.m_tiles + 1) / 2 * p.n_tiles;
    long long clusters = sm_count() / 2;
    if (p.frame_list && clusters > 16) clusters = 16;      // exact-fallback launch (idle in the steady state): a handful of CTAs, like the fallback head kernel
    if (clusters > pairs) clusters = pairs;
    kern<<<(unsigned)(2 * clusters), T_THREADS, T2Cfg::SMEM_BYTES, stream>>>(maps, p);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

template <int NT>
static int launch_tconv(const TConvMaps& maps, const TConvParams& p, cudaStream_t stream) {
    using Cfg = TConvCfg<NT>;
    auto kern = temporal_conv_kernel<NT>;
    { int rc_ = configure_kernel((const void*)kern, Cfg::SMEM_BYTES, false); if (rc_) return rc_; }
    int grid = sm_count(); if (grid > p.total_tiles) grid = p.total_tiles;
    kern<<<grid, T_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, p);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

}  // namespace vd

using namespace vd;

extern "C" int vd_temporal_conv(const void* x, void* y, int B, int T, int H, int W, int C,
                                const void* weight, const float* scale, const float* shift,
                                float slope, void* stream_) {
    return vd_temporal_conv_ex(x, y, B, T, H, W, C, weight, scale, shift, slope, VD_PREC_BF16, 0, stream_);
}

extern "C" int vd_temporal_conv_ex(const void* x, void* y, int B, int T, int H, int W, int C,
                                   const void* weight, const float* scale, const float* shift,
                                   float slope, int precision, int window_stride_frames, void* stream_) {
    return vd::temporal_conv_impl(x, y, B, T, H, W, C, weight, scale, shift, slope, precision, window_stride_frames, nullptr, nullptr, stream_);
}

int vd::temporal_conv_impl(const void* x, void* y, int B, int T, int H, int W, int C,
                           const void* weight, const float* scale, const float* shift,
                           float slope, int precision, int window_stride_frames, const unsigned int* cond, const unsigned int* frame_list, void* stream_) {
    VD_CHECK_ARG(weight && scale && shift && (B == 0 || (x && y)), "temporal_conv: null pointer");
    VD_CHECK_ARG(B >= 0 && T >= 1 && H > 0 && W > 0, "temporal_conv: bad shape");
    VD_CHECK_ARG(C >= 128 && C % 128 == 0 && C <= 1024, "temporal_conv: C = %d must be a multiple of 128, at most 1024", C);
    VD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)weight & 15) == 0, "temporal_conv: tensors must be 16-byte aligned");
    VD_CHECK_ARG(precision == VD_PREC_BF16 || precision == VD_PREC_FP32_SPLIT || precision == VD_PREC_BF16X2, "temporal_conv: precision %d", precision);
    if (B == 0) return VD_OK;
    const int planes = precision == VD_PREC_FP32_SPLIT ? 3 : (precision == VD_PREC_BF16X2 ? 2 : 1);
    // Windows sliding over a resident clip: window b starts window_stride_frames frames after window b-1 in x (0 = T: windows
    // materialised back to back).  The 3-D map (C, T*HW, B) then has OVERLAPPING windows along its last dimension; its middle
    // extent stays T*HW, so the taps that leave a window still read as zeros (TMA out-of-bounds fill) although the clip holds
    // real frames there -- the window-local zero padding of the Conv3D is kept without copying a single frame.
    const int wstride = window_stride_frames > 0 ? window_stride_frames : T;
    VD_CHECK_ARG(planes == 1 || wstride == T, "temporal_conv: the fp32-parity modes take materialised windows (window_stride_frames = T)");
    const int NT = (C % 256 == 0 && planes == 1) ? 256 : 128;      // the parity modes sum their chunks in registers: 128-wide channel blocks
    TConvParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.T = T; p.HW = H * W; p.C = C; p.rows = T * H * W;
    p.m_tiles = ceil_div(p.rows, T_BLOCK_M); p.n_tiles = C / NT;
    long long total = (long long)B * p.m_tiles * p.n_tiles;
    VD_CHECK_ARG(total < (1ll << 31), "temporal_conv: too many tiles");
    p.total_tiles = (int)total;
    p.scale = scale; p.shift = shift; p.slope = slope; p.y = (__nv_bfloat16*)y; p.cond = cond; p.frame_list = frame_list;
    p.planes = planes; p.n_prod = 1; p.y_plane_stride = (long long)B * p.rows * C;
    { static const int dbg = []() { const char* e = getenv("VD_TCONV_DBG"); return e ? atoi(e) : 0; }(); p.dbg = dbg; }
    { static const int pf = []() { const char* e = getenv("VD_TCONV_PREFETCH"); return e ? atoi(e) : 1; }(); p.prefetch = pf; }
    if (planes > 1) {                      // same plane products as the head kernel (head.cu split_products): low-order first, p0 w0 last
        static const signed char a3[6] = {0, 2, 1, 0, 1, 0}, w3[6] = {2, 0, 1, 1, 0, 0};
        static const signed char a2[3] = {0, 1, 0}, w2[3] = {1, 0, 0};
        p.n_prod = planes == 3 ? 6 : 3;
        for (int i = 0; i < p.n_prod; ++i) { p.a_pl[i] = planes == 3 ? a3[i] : a2[i]; p.w_pl[i] = planes == 3 ? w3[i] : w2[i]; }
    }
    TConvMaps maps;
    uint64_t dimsX[3] = {(uint64_t)C, (uint64_t)p.rows, (uint64_t)B * planes};
    uint64_t strX[2] = {(uint64_t)C * 2, (uint64_t)wstride * H * W * C * 2};
    uint32_t boxX[3] = {T_BLOCK_K, T_BLOCK_M, 1};
    int rc = encode_tmap_bf16(&maps.x, x, 3, dimsX, strX, boxX);
    if (rc) return rc;
    uint64_t dimsW[3] = {(uint64_t)C, (uint64_t)C, (uint64_t)3 * planes};
    uint64_t strW[2] = {(uint64_t)C * 2, (uint64_t)C * C * 2};
    static const bool pair_ok = []() { const char* e = getenv("VD_CONV_PAIR"); return e ? atoi(e) != 0 : true; }();   // CTA pairs (cta_group::2) for 256-wide channel blocks
    const bool pair = pair_ok && planes == 1 && NT == 256 && (long long)B * p.m_tiles >= 2;      // the parity modes run on the 1-CTA kernel
    uint32_t boxW[3] = {T_BLOCK_K, (uint32_t)(pair ? NT / 2 : NT), 1};
    rc = encode_tmap_bf16(&maps.w, weight, 3, dimsW, strW, boxW);
    if (rc) return rc;
    if (!pair) p.frame_list = nullptr;                     // the 1-CTA kernels recompute every window (still only when *cond != 0)
    if (pair) {
        uint64_t dimsY[3] = {(uint64_t)C, (uint64_t)p.rows, (uint64_t)B};
        uint64_t strY[2] = {(uint64_t)C * 2, (uint64_t)p.rows * C * 2};
        uint32_t boxY[3] = {64, T_BLOCK_M, 1};
        rc = encode_tmap_bf16(&maps.y, y, 3, dimsY, strY, boxY);
        if (rc) return rc;
        return launch_tconv_pair(maps, p, (cudaStream_t)stream_);
    }
    maps.y = maps.x;      // unused by the 1-CTA kernels
    if (NT == 256) return launch_tconv<256>(maps, p, (cudaStream_t)stream_);
    return launch_tconv<128>(maps, p, (cudaStream_t)stream_);
}
