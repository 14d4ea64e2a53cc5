// vd_conv_bn_lrelu -- the conv-BN-LeakyReLU cells of YOLODetectionBlockV3:
//   _conv2d(channel, k, pad, 1)            models/definitions/layers.py:63-70   (k = 1 or 3, pad = k/2)
//   _conv3d(channel, (kt,kh,kw), pad, 1)   models/definitions/layers.py:73-79   ((1,1,1), (1,3,3), (3,1,1), (3,3,3))
//   = Conv(no bias, stride 1, 'same' zero padding) + BatchNorm(eps 1e-5, inference) + LeakyReLU(0.1),
//   as stacked by YOLODetectionBlockV3 (yolo3_temporal.py:198-239; yolo3.py twin): 1x1 reduce / 3x3 expand x2,
//   1x1 reduce (= route), 3x3 tip.
//
// One implicit-GEMM kernel on tcgen05 for every kernel shape: activations are channels-last bf16
// (B, T, H, W, Cin) seen through ONE 5-D TMA map (Cin, W, H, F1, F2) -- (F1,F2) = (T,B) when the kernel has a temporal
// extent, (B*T,1) otherwise.  An M tile is a BW x BH x BF box of pixels x frames (BW*BH*BF <= 128 rows of 64 channels,
// 128-byte swizzle; small maps put the same rows of several frames into one tile: 13x13 -> 13x1x9 = 117 of 128 rows);
// tap (dt,dy,dx) of the kernel is the same box fetched at (x0+dx, y0+dy, f0+dt): the zero padding in x, y AND t is the
// TMA out-of-bounds fill, so no halo copies, no im2col buffer and no border branches exist.  Taps whose box is entirely
// outside the frame / window are skipped.
// 1x1x1 convs flatten (B,T,H,W) into one row axis (no padding needed => full 128-row tiles).
//   warp 0: TMA producer (A box + W_tap tile [NT x 64])          warp 1: MMA issuer (M128 x NT x K16, 2 TMEM accumulators)
//   warps 2-9: epilogue (column halves): tcgen05.ld -> folded BN -> LeakyReLU -> bf16 -> 128-bit stores
// 256-wide channel blocks run on CTA PAIRS (conv_bn_lrelu_pair_kernel below: tcgen05 cta_group::2, M256 x N256 MMAs, each CTA
// stages only half of the weight tile); 128-wide ones on the single-CTA kernel.
#include <stdlib.h>
#include "tc.cuh"

namespace vd {

constexpr int C_BLOCK_M = 128;
constexpr int C_BLOCK_K = 64;
constexpr int C_THREADS = 320;

struct ConvParams {
    int F2, F1, H, W, Cin, Cout;     // F1 = frame axis a tile may span (and the temporal taps shift), F2 = outer batch axis
    int kt, kh, kw;
    int BW, BH, BF;                  // pixel x frame box of one M tile
    int tiles_x, tiles_y, tiles_f, n_tiles, total_tiles;
    const float* scale; const float* shift; float slope;
    __nv_bfloat16* y;
};
struct ConvMaps { CUtensorMap x; CUtensorMap w; CUtensorMap y; };      // y: output map of the TMA-store epilogues, box {64 ch, BW, BH, BF, 1}

// Epilogue of the 128-wide kernel and of the CTA-pair kernel (r02): BN / LeakyReLU / bf16 -> shared-memory tile in the 128-byte-swizzled
// layout of a TMA box -> ONE cp.async.bulk.tensor store per [tile rows x 64 ch] (out-of-frame rows are clipped by the map).  The
// per-thread 16-byte global stores it replaces (32 rows x 16 B per instruction) cost ~4 k LSU cycles per tile: more than the whole
// mainloop of a pointwise cell.  One staging tile per column half (4 epilogue warps), named barriers 1 / 2.
constexpr int C_OUT_TILE_BYTES = C_BLOCK_M * 64 * 2;
template <int NT> struct ConvCfg {
    static constexpr int A_BYTES = C_BLOCK_M * C_BLOCK_K * 2;
    static constexpr int B_BYTES = NT * C_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr bool TMA_STORE = (NT == 128);              // 256-wide 1-CTA tiles (48 KB stages) keep the direct stores; they run on CTA pairs normally
    static constexpr int OUT_BYTES = TMA_STORE ? 2 * C_OUT_TILE_BYTES : 0;
    static constexpr int STAGES = (200 * 1024 - OUT_BYTES) / STAGE_BYTES > 8 ? 8 : (200 * 1024 - OUT_BYTES) / STAGE_BYTES;
    static constexpr int BN_BYTES = 2 * 1024 * 4;
    static constexpr int TMEM_COLS = 2 * NT;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + BN_BYTES + 2048 + 1024;
};
__device__ __forceinline__ void conv_half_bar(int half) {
    if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
}
// one epilogue pass of a warp: 64 accumulator columns of its 32 rows -> BN -> LeakyReLU -> bf16 -> its 32 rows of the staging tile
__device__ __forceinline__ void conv_stage_64(const uint32_t* v, const float* sc, const float* sf, float slope, uint32_t dst_row, uint32_t sw) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = j * 8 + h * 4;
            const float4 s4 = *reinterpret_cast<const float4*>(sc + i);
            const float4 f4 = *reinterpret_cast<const float4*>(sf + i);
            float v0 = fmaf(__uint_as_float(v[i]), s4.x, f4.x), v1 = fmaf(__uint_as_float(v[i + 1]), s4.y, f4.y);
            float v2 = fmaf(__uint_as_float(v[i + 2]), s4.z, f4.z), v3 = fmaf(__uint_as_float(v[i + 3]), s4.w, f4.w);
            v0 = v0 > 0.f ? v0 : v0 * slope; v1 = v1 > 0.f ? v1 : v1 * slope;
            v2 = v2 > 0.f ? v2 : v2 * slope; v3 = v3 > 0.f ? v3 : v3 * slope;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
            pk[2 * h] = *reinterpret_cast<uint32_t*>(&h0); pk[2 * h + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_row + (((uint32_t)j ^ sw) << 4)),
                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
    }
}

struct ConvShared {
    uint64_t full[8], empty[8], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

struct ConvTile { int b, f0, y0, x0, nt; };

template <int NT>
__global__ void __launch_bounds__(C_THREADS, 1)
conv_bn_lrelu_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<NT>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;
    unsigned char* outb = smem + Cfg::STAGES * Cfg::STAGE_BYTES;          // [half][tile rows x 128 B] staging of the TMA-store epilogue
    float* sscale = reinterpret_cast<float*>(outb + Cfg::OUT_BYTES);
    float* sshift = sscale + 1024;
    ConvShared* sh = reinterpret_cast<ConvShared*>(outb + Cfg::OUT_BYTES + Cfg::BN_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < p.Cout; i += C_THREADS) { sscale[i] = p.scale[i]; sshift[i] = p.shift[i]; }
    // rows of the A stages that no TMA box ever writes (BW*BH < 128) must still hold finite bf16 values: zero them once
    for (int i = threadIdx.x; i < Cfg::STAGES * Cfg::STAGE_BYTES / 16; i += C_THREADS)
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) { tc::mbar_init(&sh->full[i], 1); tc::mbar_init(&sh->empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&sh->tmem_full[i], 1); tc::mbar_init(&sh->tmem_empty[i], 8); }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&maps.x); tc::prefetch_tmap(&maps.w); tc::prefetch_tmap(&maps.y);
    }
    if (warp == 1) tc::tmem_alloc<Cfg::TMEM_COLS>(&sh->tmem_base);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy zero fill before async-proxy (TMA/UMMA) use
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;
    const int kb_per_tap = p.Cin / C_BLOCK_K;
    const int ntaps = p.kt * p.kh * p.kw;
    const uint32_t a_tx_bytes = (uint32_t)(p.BW * p.BH * p.BF) * C_BLOCK_K * 2;

    // tile -> (window b, frame t, box origin, channel block); channel block fastest so the A boxes are re-read from L2
    auto coords = [&](int tile) {
        ConvTile c;
        c.nt = tile % p.n_tiles; int r = tile / p.n_tiles;
        c.x0 = (r % p.tiles_x) * p.BW; r /= p.tiles_x;
        c.y0 = (r % p.tiles_y) * p.BH; r /= p.tiles_y;
        c.f0 = (r % p.tiles_f) * p.BF; c.b = r / p.tiles_f;
        return c;
    };
    auto tap_offsets = [&](int tap, int& dt, int& dy, int& dx) {
        dx = tap % p.kw - (p.kw >> 1); int r = tap / p.kw;
        dy = r % p.kh - (p.kh >> 1); dt = r / p.kh - (p.kt >> 1);
    };
    auto tap_active = [&](const ConvTile& c, int dt, int dy, int dx) -> bool {   // does the shifted box touch the frame at all?
        const int x1 = min(c.x0 + p.BW, p.W) - 1, y1 = min(c.y0 + p.BH, p.H) - 1, f1 = min(c.f0 + p.BF, p.F1) - 1;
        return (f1 + dt >= 0) && (c.f0 + dt < p.F1) && (y1 + dy >= 0) && (c.y0 + dy < p.H) && (x1 + dx >= 0) && (c.x0 + dx < p.W);
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const ConvTile c = coords(tile);
                for (int tap = 0; tap < ntaps; ++tap) {
                    int dt, dy, dx; tap_offsets(tap, dt, dy, dx);
                    if (!tap_active(c, dt, dy, dx)) continue;
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        tc::mbar_wait(&sh->empty[stage], phase ^ 1u);
                        unsigned char* a_dst = ring + stage * Cfg::STAGE_BYTES;
                        tc::mbar_expect_tx(&sh->full[stage], a_tx_bytes + Cfg::B_BYTES);
                        tc::tma_load_5d(a_dst, &maps.x, &sh->full[stage], kb * C_BLOCK_K, c.x0 + dx, c.y0 + dy, c.f0 + dt, c.b);
                        tc::tma_load_3d(a_dst + Cfg::A_BYTES, &maps.w, &sh->full[stage], kb * C_BLOCK_K, c.nt * NT, tap);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(C_BLOCK_M, NT);
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const ConvTile c = coords(tile);
                const uint32_t buf = it & 1u;
                tc::mbar_wait(&sh->tmem_empty[buf], ((it >> 1) & 1u) ^ 1u);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * NT;
                uint32_t first = 1;
                for (int tap = 0; tap < ntaps; ++tap) {
                    int dt, dy, dx; tap_offsets(tap, dt, dy, dx);
                    if (!tap_active(c, dt, dy, dx)) continue;
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        tc::mbar_wait(&sh->full[stage], phase);
                        tc::fence_after_sync();
                        const uint32_t a_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
                        const uint64_t da = tc::make_smem_desc_sw128(a_addr);
                        const uint64_t db = tc::make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < C_BLOCK_K / 16; ++k) {
                            tc::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
                            first = 0;
                        }
                        tc::umma_commit(&sh->empty[stage]);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
                tc::umma_commit(&sh->tmem_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;                               // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                     // which half of the tile's columns
        constexpr int NH = NT / 2;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int r = q * 32 + lane;                          // row of the tile = (frame lf, pixel ly, lx) of the box
        const int lx = r % p.BW, ly = (r / p.BW) % p.BH, lf = r / (p.BW * p.BH);
        uint32_t it = 0;
        if constexpr (Cfg::TMA_STORE) {
            // NT == 128: each half's 4 warps own 64 columns = one staging tile and one TMA store per M tile
            const bool issuer = (q == 0 && lane == 0);
            unsigned char* obase = outb + half * C_OUT_TILE_BYTES;
            const uint32_t orow = tc::smem_u32(obase) + (uint32_t)r * 128u;
            const uint32_t sw = (uint32_t)(r & 7);
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const ConvTile c = coords(tile);
                const uint32_t buf = it & 1u;
                tc::mbar_wait(&sh->tmem_full[buf], (it >> 1) & 1u);
                tc::fence_after_sync();
                const uint32_t tb = tmem_base + buf * NT + (uint32_t)(half * NH) + lane_addr;
                uint32_t v[64];
                tc::tmem_ld16(tb, v); tc::tmem_ld16(tb + 16, v + 16); tc::tmem_ld16(tb + 32, v + 32); tc::tmem_ld16(tb + 48, v + 48);
                tc::tmem_ld_wait();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
                const int col0 = c.nt * NT + half * NH;
                if (issuer) tc::tma_store_wait_read<0>();             // the previous tile's store has read the staging tile
                __syncwarp();
                conv_half_bar(half);
                conv_stage_64(v, sscale + col0, sshift + col0, p.slope, orow, sw);
                tc::fence_proxy_async_smem();
                __syncwarp();
                conv_half_bar(half);
                if (issuer) { tc::tma_store_5d(&maps.y, obase, col0, c.x0, c.y0, c.f0, c.b); tc::tma_store_commit(); }
            }
            if (issuer) tc::tma_store_wait<0>();
            it = 0xffffffffu;
        }
        for (int tile = blockIdx.x; tile < p.total_tiles && it != 0xffffffffu; tile += gridDim.x, ++it) {
            const ConvTile c = coords(tile);
            const uint32_t buf = it & 1u;
            const int x = c.x0 + lx, y = c.y0 + ly;
            const int f = c.f0 + lf;
            const bool inb = (lf < p.BF) && (x < p.W) && (y < p.H) && (f < p.F1);
            tc::mbar_wait(&sh->tmem_full[buf], (it >> 1) & 1u);
            tc::fence_after_sync();
            const uint32_t tbase = tmem_base + buf * NT + (uint32_t)(half * NH) + lane_addr;
            const int col0 = c.nt * NT + half * NH;
            const size_t pix = inb ? (((size_t)c.b * p.F1 + f) * p.H + y) * p.W + x : 0;
            __nv_bfloat16* yrow = p.y + pix * p.Cout + col0;
            const float* sc = sscale + col0;
            const float* sf = sshift + col0;
#pragma unroll 1
            for (int n0 = 0; n0 < NH; n0 += 32) {
                uint32_t v[32];
                tc::tmem_ld16(tbase + n0, v); tc::tmem_ld16(tbase + n0 + 16, v + 16); tc::tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 s4 = *reinterpret_cast<const float4*>(sc + n0 + i);
                    const float4 f4 = *reinterpret_cast<const float4*>(sf + n0 + i);
                    float v0 = fmaf(__uint_as_float(v[i]), s4.x, f4.x), v1 = fmaf(__uint_as_float(v[i + 1]), s4.y, f4.y);
                    float v2 = fmaf(__uint_as_float(v[i + 2]), s4.z, f4.z), v3 = fmaf(__uint_as_float(v[i + 3]), s4.w, f4.w);
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
    __syncwarp();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int NT>
static int launch_conv(const ConvMaps& maps, const ConvParams& p, cudaStream_t stream) {
    using Cfg = ConvCfg<NT>;
    auto kern = conv_bn_lrelu_kernel<NT>;
    { int rc_ = configure_kernel((const void*)kern, Cfg::SMEM_BYTES, false); if (rc_) return rc_; }
    int grid = sm_count(); if (grid > p.total_tiles) grid = p.total_tiles;
    kern<<<grid, C_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, p);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster (one TPC) work on two M tiles x the same 256 output
// channels with ONE M256 x N256 x K16 MMA stream issued by the leader.  Each CTA stages its own A box (128 rows) and only
// HALF of the weight tile (128 of the 256 rows), so the operand bytes a CTA pulls from L2 per k-block drop from 48 KB to
// 32 KB -- the 1-CTA kernel is bound by exactly that traffic.  Protocol (every barrier lives at the same offset in both CTAs):
//   producer (warp 0, both CTAs)   waits its own empty[s], TMA-loads its A box + its half of W into its own shared memory; the
//                                  loads of both CTAs report their bytes to the LEADER's full[s] (cp.async.bulk.tensor.cta_group::2)
//   MMA      (warp 1, leader CTA)  waits full[s], issues the cta_group::2 MMAs, commit (multicast) -> empty[s] of
//                                  both CTAs; after the tile's last k-block commit (multicast) -> tmem_full[buf] of both CTAs
//   epilogue (warps 2-9, both)     drains its own CTA's accumulator (its M tile), arrives on the LEADER's tmem_empty[buf]
// ---------------------------------------------------------------------------------------------------------------
struct Conv2Shared {
    uint64_t full[8], empty[8], peer_full[8], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};
template <int NT> struct Conv2Cfg {
    static constexpr int A_BYTES = C_BLOCK_M * C_BLOCK_K * 2;
    static constexpr int B_BYTES = (NT / 2) * C_BLOCK_K * 2;          // this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_BYTES = 2 * C_OUT_TILE_BYTES;            // one staging tile per column half (TMA-store epilogue)
    static constexpr int STAGES = (212 * 1024 - OUT_BYTES) / STAGE_BYTES > 8 ? 8 : (212 * 1024 - OUT_BYTES) / STAGE_BYTES;
    static constexpr int BN_BYTES = 2 * 1024 * 4;
    static constexpr int TMEM_COLS = 2 * NT;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + BN_BYTES + 2048 + 1024;
};

template <int NT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C_THREADS, 1)
conv_bn_lrelu_pair_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvParams p) {
    using Cfg = Conv2Cfg<NT>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;
    unsigned char* outb = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    float* sscale = reinterpret_cast<float*>(outb + Cfg::OUT_BYTES);
    float* sshift = sscale + 1024;
    Conv2Shared* sh = reinterpret_cast<Conv2Shared*>(outb + Cfg::OUT_BYTES + Cfg::BN_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    for (int i = threadIdx.x; i < p.Cout; i += C_THREADS) { sscale[i] = p.scale[i]; sshift[i] = p.shift[i]; }
    for (int i = threadIdx.x; i < Cfg::STAGES * Cfg::STAGE_BYTES / 16; i += C_THREADS)
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) { tc::mbar_init(&sh->full[i], 1); tc::mbar_init(&sh->empty[i], 1); tc::mbar_init(&sh->peer_full[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&sh->tmem_full[i], 1); tc::mbar_init(&sh->tmem_empty[i], 16); }   // 8 epilogue warps x 2 CTAs
        tc::fence_barrier_init();
        tc::prefetch_tmap(&maps.x); tc::prefetch_tmap(&maps.w); tc::prefetch_tmap(&maps.y);
    }
    if (warp == 1) tc::tmem_alloc_2cta<Cfg::TMEM_COLS>(&sh->tmem_base);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();                 // the peer's barriers are initialised before anything signals them
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;
    const int kb_per_tap = p.Cin / C_BLOCK_K;
    const int ntaps = p.kt * p.kh * p.kw;
    const uint32_t a_tx_bytes = (uint32_t)(p.BW * p.BH * p.BF) * C_BLOCK_K * 2;
    const int m_tiles = p.F2 * p.tiles_f * p.tiles_y * p.tiles_x;
    const int m_pairs = (m_tiles + 1) >> 1;
    const int total_pairs = m_pairs * p.n_tiles;

    // pair -> (M tile of CTA r, channel block); a pair index past the last M tile yields b >= F2: an all-padding tile
    auto coords = [&](int pair, uint32_t r) {
        ConvTile c;
        c.nt = pair % p.n_tiles;
        int m = (pair / p.n_tiles) * 2 + (int)r;
        c.x0 = (m % p.tiles_x) * p.BW; m /= p.tiles_x;
        c.y0 = (m % p.tiles_y) * p.BH; m /= p.tiles_y;
        c.f0 = (m % p.tiles_f) * p.BF; c.b = m / p.tiles_f;
        return c;
    };
    auto tap_offsets = [&](int tap, int& dt, int& dy, int& dx) {
        dx = tap % p.kw - (p.kw >> 1); int r = tap / p.kw;
        dy = r % p.kh - (p.kh >> 1); dt = r / p.kh - (p.kt >> 1);
    };
    auto tap_active1 = [&](const ConvTile& c, int dt, int dy, int dx) -> bool {
        const int x1 = min(c.x0 + p.BW, p.W) - 1, y1 = min(c.y0 + p.BH, p.H) - 1, f1 = min(c.f0 + p.BF, p.F1) - 1;
        return (c.b < p.F2) && (f1 + dt >= 0) && (c.f0 + dt < p.F1) && (y1 + dy >= 0) && (c.y0 + dy < p.H) && (x1 + dx >= 0) && (c.x0 + dx < p.W);
    };
    // both CTAs walk the same k-loop: a tap is skipped only if it is pure padding for BOTH M tiles of the pair
    auto tap_active = [&](int pair, int tap) -> bool {
        int dt, dy, dx; tap_offsets(tap, dt, dy, dx);
        return tap_active1(coords(pair, 0), dt, dy, dx) || tap_active1(coords(pair, 1), dt, dy, dx);
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
                const ConvTile c = coords(pair, rank);
                for (int tap = 0; tap < ntaps; ++tap) {
                    if (!tap_active(pair, tap)) continue;
                    int dt, dy, dx; tap_offsets(tap, dt, dy, dx);
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        tc::mbar_wait_cluster(&sh->empty[stage], phase ^ 1u);
                        unsigned char* a_dst = ring + stage * Cfg::STAGE_BYTES;
                        // both CTAs' loads report to the LEADER's full[stage]; the leader expects the bytes of the pair
                        if (rank == 0) tc::mbar_expect_tx(&sh->full[stage], 2u * (a_tx_bytes + Cfg::B_BYTES));
                        const uint32_t bar = tc::mapa_u32(&sh->full[stage], 0u);
                        tc::tma_load_5d_pair(a_dst, &maps.x, bar, kb * C_BLOCK_K, c.x0 + dx, c.y0 + dy, c.f0 + dt, c.b);
                        tc::tma_load_3d_pair(a_dst + Cfg::A_BYTES, &maps.w, bar, kb * C_BLOCK_K, c.nt * NT + (int)rank * (NT / 2), tap);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            if (rank == 0) {
                constexpr uint32_t idesc = tc::make_idesc_bf16(2 * C_BLOCK_M, NT);
                for (int pair = cluster_id; pair < total_pairs; pair += num_clusters, ++it) {
                    const uint32_t buf = it & 1u;
                    tc::mbar_wait_cluster(&sh->tmem_empty[buf], ((it >> 1) & 1u) ^ 1u);
                    tc::fence_after_sync();
                    const uint32_t d_tmem = tmem_base + buf * NT;
                    uint32_t first = 1;
                    for (int tap = 0; tap < ntaps; ++tap) {
                        if (!tap_active(pair, tap)) continue;
                        for (int kb = 0; kb < kb_per_tap; ++kb) {
                            tc::mbar_wait_cluster(&sh->full[stage], phase);
                            tc::fence_after_sync();
                            const uint32_t a_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
                            const uint64_t da = tc::make_smem_desc_sw128(a_addr);
                            const uint64_t db = tc::make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
                            for (int k = 0; k < C_BLOCK_K / 16; ++k) {
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
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int NH = NT / 2;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int r = q * 32 + lane;
        const bool issuer = (q == 0 && lane == 0);
        unsigned char* obase = outb + half * C_OUT_TILE_BYTES;
        const uint32_t orow = tc::smem_u32(obase) + (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        uint32_t it = 0;
        for (int pair = cluster_id; pair < total_pairs; pair += num_clusters, ++it) {
            const ConvTile c = coords(pair, rank);
            const uint32_t buf = it & 1u;
            tc::mbar_wait_cluster(&sh->tmem_full[buf], (it >> 1) & 1u);
            tc::fence_after_sync();
            const uint32_t tbase = tmem_base + buf * NT + (uint32_t)(half * NH) + lane_addr;
            const int col0 = c.nt * NT + half * NH;
#pragma unroll 1
            for (int pass = 0; pass < NH / 64; ++pass) {
                uint32_t v[64];
                const uint32_t ta = tbase + (uint32_t)(pass * 64);
                tc::tmem_ld16(ta, v); tc::tmem_ld16(ta + 16, v + 16); tc::tmem_ld16(ta + 32, v + 32); tc::tmem_ld16(ta + 48, v + 48);
                tc::tmem_ld_wait();
                if (pass == NH / 64 - 1) {                           // this warp has read its share of the accumulator
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive_cluster(&sh->tmem_empty[buf], 0u);
                }
                if (issuer) tc::tma_store_wait_read<0>();             // the previous store has read the staging tile
                __syncwarp();
                conv_half_bar(half);
                conv_stage_64(v, sscale + col0 + pass * 64, sshift + col0 + pass * 64, p.slope, orow, sw);
                tc::fence_proxy_async_smem();
                __syncwarp();
                conv_half_bar(half);
                if (issuer && c.b < p.F2) { tc::tma_store_5d(&maps.y, obase, col0 + pass * 64, c.x0, c.y0, c.f0, c.b); tc::tma_store_commit(); }
            }
        }
        if (issuer) tc::tma_store_wait<0>();
    }
    __syncwarp();
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();                 // neither CTA may exit (or free TMEM) while its peer can still touch it
    if (warp == 1) tc::tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
}

static int launch_conv_pair(const ConvMaps& maps, const ConvParams& p, cudaStream_t stream) {
    using Cfg = Conv2Cfg<256>;
    auto kern = conv_bn_lrelu_pair_kernel<256>;
    { int rc_ = configure_kernel((const void*)kern, Cfg::SMEM_BYTES, false); if (rc_) return rc_; }
    const long long m_tiles = (long long)p.F2 * p.tiles_f * p.tiles_y * p.tiles_x;
    const long long pairs = (m_tiles + 1) / 2 * p.n_tiles;
    long long clusters = sm_count() / 2; if (clusters > pairs) clusters = pairs;
    kern<<<(unsigned)(2 * clusters), C_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, p);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

// Box of an M tile: the (BW, BH, BF) with BW*BH*BF <= 128 that wastes the fewest MMA rows over all F frames
// (ties: the widest, then tallest box = the longest contiguous runs in memory).
static void choose_box(int H, int W, int F, int* BW, int* BH, int* BF) {
    double best = -1.0; *BW = 1; *BH = 1; *BF = 1;
    for (int bw = 1; bw <= W && bw <= C_BLOCK_M; ++bw)
        for (int bh = 1; bh <= H && bw * bh <= C_BLOCK_M; ++bh) {
            int bf = C_BLOCK_M / (bw * bh); if (bf > F) bf = F;
            const double eff = (double)W * H * F / ((double)ceil_div(W, bw) * ceil_div(H, bh) * ceil_div(F, bf) * C_BLOCK_M);
            if (eff > best + 1e-9 || (eff > best - 1e-9 && (bw > *BW || (bw == *BW && bh > *BH)))) { best = eff; *BW = bw; *BH = bh; *BF = bf; }
        }
}

}  // namespace vd

using namespace vd;

extern "C" int vd_conv_tile_box(int H, int W, int F, int* BW, int* BH, int* BF) {
    VD_CHECK_ARG(H > 0 && W > 0 && F > 0 && BW && BH && BF, "conv_tile_box: bad arguments");
    choose_box(H, W, F, BW, BH, BF);
    return VD_OK;
}

extern "C" int vd_conv_bn_lrelu(const void* x, void* y, int B, int T, int H, int W, int Cin, int Cout,
                                int kt, int kh, int kw, const void* weight, const float* scale, const float* shift,
                                float slope, void* stream_) {
    VD_CHECK_ARG(weight && scale && shift && (B == 0 || (x && y)), "conv_bn_lrelu: null pointer");
    VD_CHECK_ARG(B >= 0 && T >= 1 && H > 0 && W > 0, "conv_bn_lrelu: bad shape");
    VD_CHECK_ARG((kt == 1 || kt == 3) && (kh == 1 || kh == 3) && (kw == 1 || kw == 3),
                 "conv_bn_lrelu: kernel (%d,%d,%d): every extent must be 1 or 3 ('same' padding, stride 1)", kt, kh, kw);
    VD_CHECK_ARG(Cin >= 64 && Cin % 64 == 0, "conv_bn_lrelu: Cin = %d must be a multiple of 64", Cin);
    VD_CHECK_ARG(Cout >= 128 && Cout % 128 == 0 && Cout <= 1024, "conv_bn_lrelu: Cout = %d must be a multiple of 128, at most 1024", Cout);
    VD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)weight & 15) == 0, "conv_bn_lrelu: tensors must be 16-byte aligned");   // (an empty batch may pass null x / y)
    if (B == 0) return VD_OK;
    int NT = (Cout % 256 == 0) ? 256 : 128;
    // pointwise cells are short (85-1352 M tiles, K <= 1024): 128-wide channel blocks balance the 148 SMs better than 256-wide
    // pairs (same-box A/B: 22.0 -> 20.3 us at s32, 25.6 -> 21.9 us at s16); VD_CONV_1X1_NT256=1 restores the wide tiles
    if (kt == 1 && kh == 1 && kw == 1 && !getenv("VD_CONV_1X1_NT256")) NT = 128;
    ConvParams p;
    p.kt = kt; p.kh = kh; p.kw = kw; p.Cin = Cin; p.Cout = Cout;
    if (kt == 1 && kh == 1 && kw == 1) {          // pointwise: one flat row axis, full 128-row tiles
        const long long rows = (long long)B * T * H * W;
        VD_CHECK_ARG(rows < (1ll << 31), "conv_bn_lrelu: too many pixels");
        p.F2 = 1; p.F1 = 1; p.H = 1; p.W = (int)rows; p.BW = C_BLOCK_M; p.BH = 1; p.BF = 1;
    } else {
        VD_CHECK_ARG((long long)B * T < (1ll << 31), "conv_bn_lrelu: too many frames");
        if (kt == 1) { p.F2 = 1; p.F1 = B * T; } else { p.F2 = B; p.F1 = T; }     // frames are independent unless taps shift t
        p.H = H; p.W = W;
        choose_box(H, W, p.F1, &p.BW, &p.BH, &p.BF);
    }
    p.tiles_x = ceil_div(p.W, p.BW); p.tiles_y = ceil_div(p.H, p.BH); p.tiles_f = ceil_div(p.F1, p.BF); p.n_tiles = Cout / NT;
    const long long total = (long long)p.F2 * p.tiles_f * p.tiles_x * p.tiles_y * p.n_tiles;
    VD_CHECK_ARG(total < (1ll << 31), "conv_bn_lrelu: too many tiles");
    p.total_tiles = (int)total;
    p.scale = scale; p.shift = shift; p.slope = slope; p.y = (__nv_bfloat16*)y;
    ConvMaps maps;
    const uint64_t e = 2;
    uint64_t dimsX[5] = {(uint64_t)Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.F1, (uint64_t)p.F2};
    uint64_t strX[4] = {Cin * e, (uint64_t)p.W * Cin * e, (uint64_t)p.H * p.W * Cin * e, (uint64_t)p.F1 * p.H * p.W * Cin * e};
    uint32_t boxX[5] = {C_BLOCK_K, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BF, 1};
    int rc = encode_tmap_bf16(&maps.x, x, 5, dimsX, strX, boxX);
    if (rc) return rc;
    const int ntaps = kt * kh * kw;
    uint64_t dimsW[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)ntaps};
    uint64_t strW[2] = {Cin * e, (uint64_t)Cout * Cin * e};
    // CTA pairs (cta_group::2) for 256-wide channel blocks; VD_CONV_PAIR=0 keeps the 1-CTA kernel
    static const bool pair_ok = []() { const char* e = getenv("VD_CONV_PAIR"); return e ? atoi(e) != 0 : true; }();
    const bool pair = pair_ok && NT == 256 && (long long)p.F2 * p.tiles_f * p.tiles_y * p.tiles_x >= 2;
    uint32_t boxW[3] = {C_BLOCK_K, (uint32_t)(pair ? NT / 2 : NT), 1};
    rc = encode_tmap_bf16(&maps.w, weight, 3, dimsW, strW, boxW);
    if (rc) return rc;
    uint64_t dimsY[5] = {(uint64_t)Cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.F1, (uint64_t)p.F2};
    uint64_t strY[4] = {Cout * e, (uint64_t)p.W * Cout * e, (uint64_t)p.H * p.W * Cout * e, (uint64_t)p.F1 * p.H * p.W * Cout * e};
    uint32_t boxY[5] = {64, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BF, 1};
    rc = encode_tmap_bf16(&maps.y, y, 5, dimsY, strY, boxY);
    if (rc) return rc;
    if (pair) return launch_conv_pair(maps, p, (cudaStream_t)stream_);
    if (NT == 256) return launch_conv<256>(maps, p, (cudaStream_t)stream_);
    return launch_conv<128>(maps, p, (cudaStream_t)stream_);
}

// ---------------------------------------------------------------------------------------------------------------
// vd_upsample_concat -- the glue between two detection blocks of YOLOV3.hybrid_forward (yolo3.py:515-519):
//   upsample = _upsample(x, stride=2)                       nearest, out[y, x] = in[y / 2, x / 2]   (layers.py:10-20)
//   x = concat(slice_like(upsample, route_now, axes=(2,3)), route_now, dim=1)
// on channels-last bf16: out (B, H2, W2, C1 + C2), channels [0, C1) = upsampled x cropped to (H2, W2), [C1, C1+C2) = route.
// One 16-byte chunk (8 channels) per thread: pure HBM streaming (the upsample reads hit L2 three times out of four).
// ---------------------------------------------------------------------------------------------------------------
namespace vd {
__global__ void __launch_bounds__(256)
upsample_concat_kernel(const uint4* __restrict__ x, const uint4* __restrict__ route, uint4* __restrict__ out,
                       long long total, int H, int W, int c1v, int H2, int W2, int c2v) {
    const int cv = c1v + c2v;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        long long pix = i / cv;
        const int xx = (int)(pix % W2); pix /= W2;
        const int yy = (int)(pix % H2);
        const long long b = pix / H2;
        uint4 v;
        if (c < c1v) v = __ldg(x + ((b * H + (yy >> 1)) * W + (xx >> 1)) * c1v + c);
        else v = __ldg(route + ((b * H2 + yy) * W2 + xx) * c2v + (c - c1v));
        out[i] = v;
    }
}
}  // namespace vd

extern "C" int vd_upsample_concat(const void* x, const void* route, void* out, int B, int H, int W, int C1,
                                  int H2, int W2, int C2, void* stream_) {
    VD_CHECK_ARG(B == 0 || (x && route && out), "upsample_concat: null pointer");
    VD_CHECK_ARG(B >= 0 && H > 0 && W > 0 && H2 > 0 && W2 > 0, "upsample_concat: bad shape");
    VD_CHECK_ARG(H2 <= 2 * H && W2 <= 2 * W, "upsample_concat: the route map (%d x %d) is larger than the upsampled one (%d x %d)", H2, W2, 2 * H, 2 * W);
    VD_CHECK_ARG(C1 > 0 && C2 > 0 && C1 % 8 == 0 && C2 % 8 == 0, "upsample_concat: channel counts (%d, %d) must be multiples of 8", C1, C2);
    VD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)route & 15) == 0 && ((uintptr_t)out & 15) == 0, "upsample_concat: tensors must be 16-byte aligned");
    if (B == 0) return VD_OK;
    const long long total = (long long)B * H2 * W2 * ((C1 + C2) / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    vd::upsample_concat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>((const uint4*)x, (const uint4*)route, (uint4*)out, total,
                                                                                    H, W, C1 / 8, H2, W2, C2 / 8);
    VD_LAUNCH_CHECK();
    return VD_OK;
}
