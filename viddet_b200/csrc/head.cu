// Fused detection head for sm_100a: 1x1 prediction conv (tcgen05 GEMM, TMEM accumulators, TMA
// operand ring) with the YOLOOutputV3 decode and an exact top-k candidate filter in the epilogue,
// followed by the per-frame top-k / NMS kernel (nms_core.cuh).
//
// Replaces, per forward call (reference, inference mode):
//   nn.Conv2D(A*(5+C), 1x1)                      yolo3.py:62,157      -> GEMM  [M=pixels, N=A*(5+C), K=Cin]
//   YOLOOutputV3 decode (~15 elementwise ops)    yolo3.py:158-197     -> epilogue math on TMEM rows
//   concat of the 3 scales                       yolo3.py:523         -> row index arithmetic
//   contrib.box_nms + slice + split              yolo3.py:526-534     -> candidate lists -> nms_final
//   late 'cat' temporal join                     yolo3.py:1134-1136   -> K loop over the K frames
//
// Kernel head_kernel<EPI, C, NPAD>: persistent (one CTA per SM), 64 + G*SPLIT*128 threads:
//   warp 0      claims tiles from a global counter (dynamic, largest tiles first) and issues TMA: A tile [128 px x 64 ch]
//               (4-D map: Cin, HW, K, frames; out-of-range pixels zero-filled) + W tile [NPAD x 64 ch] per stage, mbarrier ring
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16 M=128 N=NPAD K=16, fp32 accumulators in G TMEM buffers,
//               tcgen05.commit -> empty/full barriers
//   warps 2..   epilogue warpgroups (one TMEM lane = one pixel): tcgen05.ld of the pixel's logits, decode, then
//        EPI_SPEC    (steady state) every candidate that can reach its frame slot's threshold tau (left by the previous call's
//                    NMS kernel) is scored and appended to the FRAME's list: one compare per class logit, per-warp staging,
//                    deferred flush, no barriers; nms_spec_kernel verifies each frame (list complete <=> no overflow and
//                    >= k keys >= tau) and finishes it; frames that fail are redone by the exact pair below, on the device
//        EPI_FILTER  (exact path: fallback + first call) per-tile exact top-k superset by a pivot search on the logit
//                    prefilter, frame-level running bound from a coarse histogram -> per-tile lists -> nms_final_hist_kernel
//        EPI_DET     materialise the reference's (frames, rows, 6) detection rows
//        EPI_PRED    write the conv output (B, N, H, W) fp32
//   Wide heads (NPAD > 128: two 256-column accumulators) put SPLIT = 2 warpgroups on one accumulator (class chunks interleaved).
#include <stdlib.h>
#include <type_traits>
#include "tc.cuh"
#include "nms_core.cuh"

namespace vd {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kMaxEpiGroups = 4;              // epilogue warpgroups G (2 or 3); group g owns TMEM buffer g (tiles it%G == g)
constexpr int kEpiWarp0 = 2;                  // warp 0: TMA, warp 1: MMA, then 2 x 4 epilogue warps
constexpr int kEpiThreads = 128;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;

enum { EPI_FILTER = 0, EPI_DET = 1, EPI_PRED = 2, EPI_SPEC = 3 };
#ifndef VD_SPEC_CAP
#define VD_SPEC_CAP 1024      // r02: 1024 keys (8 KB of the NMS CTA's shared memory instead of 16) let the head kernel take 200 KB = 7 ring stages while
#endif                        // the NMS CTA of the previous launch still shares the SM: 33.5 -> 32.5 us per step; 2048 + 181 KB was r01's choice
constexpr int kSpecCap = VD_SPEC_CAP;         // candidates per frame the speculative path may emit (multiple of 512)
#ifndef VD_SPEC_G
#define VD_SPEC_G 3
#endif
#ifndef VD_SPEC_SPLIT
#define VD_SPEC_SPLIT 2
#endif
#ifndef VD_FALLBACK_CTAS
#define VD_FALLBACK_CTAS 32      // CTAs of the exact fallback head kernel (idle in the steady state; measured r1: 2 / 8 / 32 CTAs -> 39.1 / 39.1 / 39.3 us per step, all-frames-failed call 2.5 / 0.98 / 0.51 ms)
#endif
constexpr int kFallbackCtas = VD_FALLBACK_CTAS;
constexpr int kSpecStage = 64;                // per-warp staging entries per tile (double-buffered); more go straight to global memory
constexpr int kHeadSharedBytes = 2048;          // room for struct HeadShared (barriers, scheduler ring, per-group state)

// Per-frame histogram of the emitted candidates' scores: 4096 bins, 512 per octave over [2^-7, 2).
// Monotone in the key, so "the highest bin b* whose suffix count reaches k" gives a frame-level
// pivot with at most k + (population of b*) candidates above it -- no search in the NMS kernel.
constexpr int kHistBins = 4096;
constexpr int kHistCap = 1024;                // candidates the NMS kernel can hold after the pivot
__host__ __device__ __forceinline__ uint32_t hist_bin(uint32_t key_hi) {
    if (!(key_hi & 0x80000000u)) return 0u;                       // negative scores: lowest bin
    uint32_t b = (key_hi & 0x7fffffffu) >> 14;
    const uint32_t lo = (127u - 7u) << 9;
    b = b > lo ? b - lo : 0u;
    return b > (uint32_t)(kHistBins - 1) ? (uint32_t)(kHistBins - 1) : b;
}
__host__ __device__ __forceinline__ uint32_t hist_edge(uint32_t bin) {   // smallest key_hi that falls in `bin` (0: everything)
    return bin == 0u ? 0u : ((((bin + ((127u - 7u) << 9)) << 14)) | 0x80000000u);
}

struct HeadKernelParams {
    HeadGeom g;
    int frames, K_frames;
    int split;                           // fp32-parity modes: operand planes in memory (0: off; 3: hi/mid/lo; 2: hi/lo); K_frames = number of plane products
    signed char a_pl[8], w_pl[8];        // plane of A / of W that product kf of the K loop multiplies (smallest products first, p0 w0 last)
    int cin[VD_MAX_SCALES];
    int pb[VD_MAX_SCALES];               // pixel blocks per frame
    int tile_start[VD_MAX_SCALES + 1];   // cumulative tile index per processing slot
    int order[VD_MAX_SCALES];            // order[j] = scale processed j-th (mid-size tiles first: short pipeline fill)
    int tif_base[VD_MAX_SCALES];         // tile-in-frame index of the scale's first pixel block
    int tiles_per_frame, total_tiles, n_pad;
    int n_valid;                         // prediction channels actually present (<= n_pad)
    // class windows (any num_class on the compiled shapes): this launch covers classes [c_off, c_off + c_valid) of c_total; the
    // kernel's own class index c (0 .. C-1 of its template shape) is class c_off + c of the head; c >= c_valid are padding
    int c_off, c_valid, n_pass, pass;
    int prefetch;                        // tiles the producer claims ahead and prefetches into L2 (0: off)
    const float* bias[VD_MAX_SCALES];
    // EPI_FILTER
    float4* boxes; uint64_t* lists; uint32_t* counts; float valid_thresh; int k, cap;
    uint32_t* counts_hi;                 // [frames][tiles] entries at the front of each list that are >= the frame's hint_hi
    uint32_t* hint_hi;                   // [frames] key_hi below which this frame's pivot is not expected (from the previous call; any value is valid)
    unsigned long long* hints;           // [2*VD_MAX_SCALES][2] (pivot, band) warm starts, persist in the workspace across calls
    uint32_t* hist;                      // [frames][kHistBins] score histogram of the emitted candidates (left zeroed by the NMS kernel)
    uint32_t* coarse;                    // [frames][64] the same histogram at 64 fine bins per bin: tiles read it for a running lower bound of the frame's k-th score
    unsigned int* tile_counter;          // [0] next tile, [1] finished CTAs, [2] ws_magic once the workspace is in its between-calls state; null: static round-robin
    unsigned int ws_magic;               // kWsMagic mixed with this call's workspace layout (a workspace last used with another shape is treated as uninitialised)
    // EPI_SPEC (speculative frame-level threshold) and its exact fallback
    uint64_t* spec_lists;                // [frames][kSpecCap] keys of the candidates >= the call's threshold
    uint32_t* spec_cnt;                  // [frames] candidates emitted (may exceed kSpecCap = overflow); left zeroed by the NMS kernel
    uint32_t* spec_tau;                  // [frames] score threshold (float bits) each frame slot is filtered with; written by the NMS kernels for the next call (any value is valid)
    uint32_t* spec_state;                // [2] failed frames, [3] finished NMS CTAs, [4] failed frames of the last call (stats)
    const uint32_t* frame_list;          // EPI_FILTER as the fallback: frames to process = frame_list[0 .. *frame_count)
    const uint32_t* frame_count;
    long long* stamps;                   // profiling aid (VD_DEBUG_HEAD_STAMPS): clock64 per tile of CTA 0, [it][8]
    int dbg;                             // profiling aid (VD_DEBUG_SKIP_EPILOGUE=level): stop the epilogue early, results are garbage
    // EPI_DET
    float* det; long long det_rows_total;
    // EPI_PRED
    float* pred[VD_MAX_SCALES]; long long pred_frame_stride;   // floats between frames (N_total*HW)
};

struct HeadMaps { CUtensorMap a[VD_MAX_SCALES]; CUtensorMap w[VD_MAX_SCALES]; };

template <int EPI, int C, int NPAD> struct HeadCfg {
    static constexpr int CPA = (C + 15) / 16;                  // TMEM read chunks per anchor (<= 16 class logits each: 3 x 16 live registers)
    static constexpr int CH = (C + CPA - 1) / CPA;
    static constexpr int CH4 = (CH + 3) / 4 * 4;               // class-bias chunk padded for 128-bit shared loads
    static constexpr int CBIAS_BYTES = (EPI == EPI_FILTER || EPI == EPI_SPEC) ? VD_MAX_SCALES * 3 * CPA * CH4 * 4 : 0;
    static constexpr int CONF_BYTES = (EPI == EPI_FILTER) ? kMaxEpiGroups * 3 * kEpiThreads * 4 : 0;
    static constexpr int LIST_BUFS = (EPI == EPI_FILTER) ? 1 : 0;
    static constexpr int B_TILE_BYTES = NPAD * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int TMEM_STRIDE = NPAD <= 128 ? 128 : 256;
    // 3 groups (12 epilogue warps, 128 registers/thread) when three accumulators fit TMEM; else 2 groups
    static constexpr int G = (EPI == EPI_FILTER && TMEM_STRIDE == 128) ? (NPAD <= 80 ? 4 : 3) : ((EPI == EPI_SPEC && TMEM_STRIDE == 128) ? (NPAD <= 96 ? VD_SPEC_G : 2) : 2);   // 4 groups only where 80 registers suffice; EPI_SPEC: 3 groups up to 96 columns, 2 (128 registers) above (measured: VOC 39.0 / 44.2 / 39.6 us with 3 / 2 / 4 groups, VID 47.4 / 46.3 / 51.8)
    // wide heads (two 256-column accumulators fill TMEM): SPLIT warpgroups drain ONE accumulator together, each taking every
    // SPLIT-th class chunk (the epilogue, not the mainloop, bounds these heads: 240 class logits per pixel at C = 80)
    static constexpr int SPLIT = (EPI == EPI_SPEC && TMEM_STRIDE == 256) ? VD_SPEC_SPLIT : 1;
    static constexpr int THREADS = kEpiWarp0 * 32 + G * SPLIT * kEpiThreads;
#ifdef VD_SPEC_MAXREG
    static constexpr int MAXREG = (EPI == EPI_SPEC) ? VD_SPEC_MAXREG : ((THREADS > 448) ? 80 : ((THREADS > 320) ? 96 : 128));   // tuning knob
#else
    static constexpr int MAXREG = (THREADS > 448) ? 80 : ((THREADS > 320) ? 96 : 128);
#endif
    static constexpr int LIST_BYTES = (EPI == EPI_SPEC) ? G * SPLIT * 4 * 2 * kSpecStage * 8 : G * LIST_BUFS * kListCap * 8;
    static constexpr int EPI_BYTES = LIST_BYTES + CBIAS_BYTES + CONF_BYTES + VD_MAX_SCALES * NPAD * 4 + kHeadSharedBytes;
    // EPI_FILTER leaves ~45 KB of the SM's shared memory to a co-resident nms_final_hist_kernel CTA of the previous batch
    #ifndef VD_SPEC_SMEM_KB
#define VD_SPEC_SMEM_KB 200
#endif
#ifndef VD_SPEC_SMEM_KB_WIDE
#define VD_SPEC_SMEM_KB_WIDE VD_SPEC_SMEM_KB
#endif
    static constexpr int SMEM_BUDGET = ((EPI == EPI_SPEC) ? (TMEM_STRIDE == 256 ? VD_SPEC_SMEM_KB_WIDE : VD_SPEC_SMEM_KB) : ((EPI == EPI_FILTER) ? 181 : 225)) * 1024;
    static constexpr int STAGES_RAW = (SMEM_BUDGET - EPI_BYTES - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = (G * TMEM_STRIDE) <= 256 ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024;   // +1024 alignment slack
    static_assert(NPAD % 16 == 0 && NPAD >= 16 && NPAD <= 256, "NPAD");
    static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024-byte alignment");
    static_assert(STAGES >= 3, "pipeline too shallow");
};

struct EpiGroupShared {
    uint32_t cnt[3]; uint32_t cursor; uint32_t cursor2; uint32_t bound; uint32_t pcur[3];
    uint32_t chist[64];                    // this tile's contribution to the frame's coarse histogram
    uint64_t guess[2 * VD_MAX_SCALES];     // warm-start pivot per (scale, full / partial pixel block)
    uint64_t band[2 * VD_MAX_SCALES];      // running estimate of the accept band's key width
};
constexpr unsigned int kWsMagic = 0x56444231u; // 'VDB1': the workspace's counters / histogram are in their between-calls state
constexpr int kSchedSlots = 8;               // tile-id ring between the producer (which claims tiles) and the MMA / epilogue roles
struct HeadShared {
    uint64_t full[8], empty[8], tmem_full[kMaxEpiGroups], tmem_empty[kMaxEpiGroups];
    uint64_t sched_full[kSchedSlots], sched_empty[kSchedSlots];
    int sched_tile[kSchedSlots];
    uint32_t spec_wcnt[4 * kMaxEpiGroups];       // EPI_SPEC: staged keys of each epilogue warp's current tile
    uint32_t tmem_base, pad_;
    EpiGroupShared grp[kMaxEpiGroups];
};

// named barrier 1 + grp, immediate ids: a register id makes ptxas reserve all 16 hardware barriers for the CTA,
// which would keep any other kernel's CTA (the co-scheduled NMS kernel) off the SM
static_assert(sizeof(HeadShared) <= kHeadSharedBytes, "HeadShared outgrew its shared-memory reservation");
__device__ __forceinline__ void epi_bar(int grp) {
    switch (grp) {
        case 0: asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory"); break;
        case 2: asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory"); break;
        default: asm volatile("bar.sync 4, %0;" ::"n"(kEpiThreads) : "memory"); break;
    }
}

// block_sum over the 128 threads of one epilogue group (named barrier 1+grp), see select.cuh::block_sum
__device__ __forceinline__ uint32_t epi_sum(uint32_t v, EpiGroupShared* s, int grp, int et, int& it) {
    uint32_t w = __reduce_add_sync(0xffffffffu, v);
    const int slot = it % 3;
    if ((et & 31) == 0 && w) atomicAdd(&s->cnt[slot], w);
    if (et == 0) s->cnt[(it + 1) % 3] = 0;
    epi_bar(grp);
    uint32_t r = s->cnt[slot];
    ++it;
    return r;
}

// fallback mode: tiles enumerate the listed frames one after another
__device__ __forceinline__ void tile_coords_list(const HeadKernelParams& p, int tile, int& s, int& f, int& pblk) {
    const int fi = tile / p.tiles_per_frame, t = tile - fi * p.tiles_per_frame;
    f = (int)p.frame_list[fi];
    s = 0;
#pragma unroll
    for (int i = 1; i < VD_MAX_SCALES; ++i) if (i < p.g.num_scales && t >= p.tif_base[i]) s = i;
    pblk = t - p.tif_base[s];
}
__device__ __forceinline__ void tile_coords(const HeadKernelParams& p, int tile, int& s, int& f, int& pblk) {
    if (p.frame_list) { tile_coords_list(p, tile, s, f, pblk); return; }
    int j = 0;
#pragma unroll
    for (int i = 1; i < VD_MAX_SCALES; ++i) if (i < p.g.num_scales && tile >= p.tile_start[i]) j = i;
    s = p.order[j];
    int local = tile - p.tile_start[j];
    f = local / p.pb[s];
    pblk = local - f * p.pb[s];
}

// Register cap: the register file is split per SM sub-partition (16K registers = 512 per lane); 4 head warps x 96 + 2 NMS
// warps x 56 = 496 lets a CTA of the co-scheduled NMS kernel share each sub-partition (3 x 128 + 112 for the 10-warp configs).
template <int EPI, int C, int NPAD>
__global__ void __maxnreg__((HeadCfg<EPI, C, NPAD>::MAXREG))
head_kernel(const __grid_constant__ HeadMaps maps, const __grid_constant__ HeadKernelParams p) {
    using Cfg = HeadCfg<EPI, C, NPAD>;
    constexpr int P = 5 + C;
    constexpr int G = Cfg::G;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment for the 128B-swizzled operand tiles; offset arithmetic keeps the
    // shared-memory address space visible to the compiler (LDS/STS instead of generic LD/ST)
    unsigned char* smem = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* ring = smem;
    uint64_t* slist = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);     // [group][buf][kListCap]
    float* scbias = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::LIST_BYTES);  // [3 scales][3 anchors][CPA][CH4] class biases, 16-byte aligned chunks
    float* sconf = scbias + Cfg::CBIAS_BYTES / 4;                                                       // [group][128 px][3] objectness of the tile being filtered
    float* sbias = sconf + Cfg::CONF_BYTES / 4;                                                         // [3][NPAD]
    (void)scbias; (void)sconf;
    HeadShared* sh = reinterpret_cast<HeadShared*>(sbias + VD_MAX_SCALES * NPAD);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the exact fallback of the speculative path with no failed frame (the steady state): nothing to set up, nothing to claim
    if constexpr (EPI == EPI_FILTER) { if (p.frame_list != nullptr && *p.frame_count == 0u) return; }
    auto gtime = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); return (long long)t; };
    if ((EPI == EPI_FILTER || EPI == EPI_SPEC) && p.stamps && p.frame_list == nullptr && threadIdx.x == 0) {
        p.stamps[4096 + blockIdx.x * 4 + 0] = gtime();
        unsigned int sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.stamps[4096 + blockIdx.x * 4 + 3] = (long long)sm;
    }

    // Programmatic dependent launch: the next kernel of the stream (the head kernel of the next batch in the pipeline)
    // may be scheduled as soon as this grid has started, so its CTAs take over SMs one by one as ours retire and run
    // their setup early.  Everything that could depend on an earlier kernel comes after griddepcontrol.wait.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if constexpr (EPI != EPI_SPEC) asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---------------- one-time setup
    // EPI_SPEC: the producer warp does not stage biases and does not wait for the rest of the CTA: it initialises the barriers,
    // signals the named barrier (bar.arrive, non-blocking) and starts claiming tiles / issuing TMA at once, so the first loads'
    // latency overlaps the bias staging and the TMEM allocation of the other warps (kEarly).
    constexpr bool kEarly = (EPI == EPI_SPEC);
    const int st0 = kEarly ? (int)threadIdx.x - 32 : (int)threadIdx.x;        // staging index of this thread (warp 0 excluded when early)
    constexpr int kStageThreads = kEarly ? Cfg::THREADS - 32 : Cfg::THREADS;
    for (int i = st0; i >= 0 && i < VD_MAX_SCALES * NPAD; i += kStageThreads) {
        int s = i / NPAD, n = i % NPAD;
        sbias[i] = (s < p.g.num_scales && p.bias[s] && n < p.n_valid) ? p.bias[s][n] : 0.0f;
    }
    if constexpr (EPI == EPI_FILTER || EPI == EPI_SPEC) {
        constexpr int P = 5 + C;
        for (int i = st0; i >= 0 && i < VD_MAX_SCALES * 3 * Cfg::CPA * Cfg::CH4; i += kStageThreads) {
            const int s = i / (3 * Cfg::CPA * Cfg::CH4), r = i % (3 * Cfg::CPA * Cfg::CH4);
            const int a = r / (Cfg::CPA * Cfg::CH4), cc = (r / Cfg::CH4) % Cfg::CPA, ci = r % Cfg::CH4;
            const int c = cc * Cfg::CH + ci;
            scbias[i] = (s < p.g.num_scales && p.bias[s] && ci < Cfg::CH && c < C) ? p.bias[s][a * P + 5 + c] : 0.0f;
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) { tc::mbar_init(&sh->full[i], 1); tc::mbar_init(&sh->empty[i], 1); }
        for (int i = 0; i < G; ++i) { tc::mbar_init(&sh->tmem_full[i], 1); tc::mbar_init(&sh->tmem_empty[i], 4 * Cfg::SPLIT); }
        for (int i = 0; i < kSchedSlots; ++i) { tc::mbar_init(&sh->sched_full[i], 1); tc::mbar_init(&sh->sched_empty[i], 1 + 4 * Cfg::SPLIT); }   // consumers: MMA thread + the epilogue warps of the tile's accumulator
        for (int i = 0; i < 4 * kMaxEpiGroups; ++i) sh->spec_wcnt[i] = 0u;
        for (int g = 0; g < G; ++g) {
            sh->grp[g].cnt[0] = sh->grp[g].cnt[1] = sh->grp[g].cnt[2] = 0; sh->grp[g].cursor = 0; sh->grp[g].cursor2 = 0; sh->grp[g].pcur[0] = sh->grp[g].pcur[1] = sh->grp[g].pcur[2] = 0;
            for (int i = 0; i < 2 * VD_MAX_SCALES; ++i) {
                unsigned long long hg = 0ull, hb = 0ull;
                if (EPI == EPI_FILTER && p.hints) { hg = p.hints[2 * i]; hb = p.hints[2 * i + 1]; }
                sh->grp[g].guess[i] = hg; sh->grp[g].band[i] = hb ? hb : (1ull << 50);
            }
        }
        tc::fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.g.num_scales; ++s) { tc::prefetch_tmap(&maps.a[s]); tc::prefetch_tmap(&maps.w[s]); }
    }
    if (warp == 1) { if (p.split) tc::tmem_alloc<512>(&sh->tmem_base); else tc::tmem_alloc<Cfg::TMEM_COLS>(&sh->tmem_base); }   // parity modes: all of TMEM for partial accumulators
    tc::fence_before_sync();
    if constexpr (kEarly) {
        if (warp == 0) { __syncwarp(); asm volatile("bar.arrive 1, %0;" ::"n"(Cfg::THREADS) : "memory"); }
        else asm volatile("bar.sync 1, %0;" ::"n"(Cfg::THREADS) : "memory");
    } else
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;
    if constexpr (EPI == EPI_SPEC) asm volatile("griddepcontrol.wait;" ::: "memory");      // setup (weights only) overlapped the previous kernel's tail
    if ((EPI == EPI_FILTER || EPI == EPI_SPEC) && p.stamps && p.frame_list == nullptr && threadIdx.x == 0) p.stamps[4096 + blockIdx.x * 4 + 1] = gtime();

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (tc::elect_one()) {
            int stage = 0; uint32_t phase = 0;
            const uint64_t pol_a = tc::policy_evict_first();    // activations are read exactly once
            const uint64_t pol_w = tc::policy_evict_last();     // weights are re-read by every tile
            // claims tiles (global counter: dynamic load balance, largest tiles first) and publishes them to the other roles
            const bool dyn = p.tile_counter != nullptr && p.tile_counter[2] == p.ws_magic;   // uninitialised workspace: static round-robin
            const uint32_t total_tiles = p.frame_list ? *p.frame_count * (uint32_t)p.tiles_per_frame : (uint32_t)p.total_tiles;
            auto claim = [&](uint32_t i) -> int {
                const uint32_t t = dyn ? atomicAdd(p.tile_counter, 1u) : (uint32_t)blockIdx.x + i * gridDim.x;
                return t < total_tiles ? (int)t : -1;
            };
            // Tiles are claimed D + 1 ahead; the activations of the farthest one are prefetched into L2 (cp.async.bulk.prefetch.tensor):
            // the shared-memory ring bounds the bytes in flight per SM (7 x 16 KB), the L2 prefetch of whole tiles lifts that bound
            // without shared memory.  D = 0: claim one ahead, no prefetch (r01 behaviour).
            const int D = (p.prefetch > 0 && !p.split) ? (p.prefetch < 3 ? p.prefetch : 3) : 0;
            auto prefetch_tile = [&](int t) {
                if (t < 0) return;
                int s2, f2, pb2; tile_coords(p, t, s2, f2, pb2);
                for (int kf2 = 0; kf2 < p.K_frames; ++kf2)
                    for (int c2 = 0; c2 < p.cin[s2]; c2 += BLOCK_K) tc::tma_prefetch_4d(&maps.a[s2], c2, pb2 * BLOCK_M, kf2, f2);
            };
            uint32_t nclaim = 0u;
            const int QN = D > 1 ? D : 1;                           // tiles held at the top of an iteration: q[0] = the next to run
            int q[4] = {-1, -1, -1, -1};
            for (int i = 0; i < QN; ++i) {
                q[i] = (i == 0 || q[i - 1] >= 0) ? claim(nclaim++) : -1;
                if (D > 0 && i > 0) prefetch_tile(q[i]);
            }
            int n_end = 0;
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it % kSchedSlots;
                tc::mbar_wait(&sh->sched_empty[slot], ((it / kSchedSlots) & 1u) ^ 1u);
                const int tile = q[0];
                if (tile >= 0) {                                    // shift the queue; the new claim is in flight while this tile's loads are issued
                    for (int i = 0; i + 1 < QN; ++i) q[i] = q[i + 1];
                    q[QN - 1] = claim(nclaim++);
                    if (D > 0) prefetch_tile(q[QN - 1]);
                }
                sh->sched_tile[slot] = tile;
                tc::mbar_arrive(&sh->sched_full[slot]);
                if (tile < 0) { if (++n_end == G) break; continue; }   // one terminator per epilogue group
                int s, f, pblk; tile_coords(p, tile, s, f, pblk);
                const int kb_per_frame = p.cin[s] / BLOCK_K;
                const int nkb = kb_per_frame * p.K_frames;
                int kf = 0, c0 = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    tc::mbar_wait(&sh->empty[stage], phase ^ 1u);
                    unsigned char* a_dst = ring + stage * Cfg::STAGE_BYTES;
                    unsigned char* b_dst = a_dst + A_TILE_BYTES;
                    tc::mbar_expect_tx(&sh->full[stage], Cfg::STAGE_BYTES);
                    // fp32-parity modes: product kf of the K loop pairs plane a_pl[kf] of A with plane w_pl[kf] of W (split_products())
                    const int a_pl = p.split ? (int)p.a_pl[kf & 7] : kf;
                    const int w_pl = p.split ? (int)p.w_pl[kf & 7] : kf;
                    tc::tma_load_4d_hint(a_dst, &maps.a[s], &sh->full[stage], c0, pblk * BLOCK_M, a_pl, f, pol_a);
                    tc::tma_load_2d_hint(b_dst, &maps.w[s], &sh->full[stage], w_pl * p.cin[s] + c0, 0, pol_w);
                    c0 += BLOCK_K; if (c0 == p.cin[s]) { c0 = 0; ++kf; }
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(BLOCK_M, NPAD);
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            for (;; ++it) {
                const uint32_t slot = it % kSchedSlots;
                tc::mbar_wait(&sh->sched_full[slot], (it / kSchedSlots) & 1u);
                const int tile = sh->sched_tile[slot];
                tc::mbar_arrive(&sh->sched_empty[slot]);
                if (tile < 0) break;
                int s, f, pblk; tile_coords(p, tile, s, f, pblk);
                const int nkb = (p.cin[s] / BLOCK_K) * p.K_frames;
                const uint32_t buf = it % (uint32_t)G;
                if (p.stamps && blockIdx.x == 0 && it < 250u) p.stamps[it * 16 + 0] = clock64();
                tc::mbar_wait(&sh->tmem_empty[buf], ((it / (uint32_t)G) & 1u) ^ 1u);
                tc::fence_after_sync();
                if (p.stamps && blockIdx.x == 0 && it < 250u) p.stamps[it * 16 + 1] = clock64();
                const uint32_t d_tmem = tmem_base + buf * Cfg::TMEM_STRIDE;
                if (p.split) {
                    // fp32-parity modes: partial accumulators (see split_products): 0 = all low-order products, 1.. = K-ranges of p0 w0;
                    // the tile owns every TMEM buffer, so it starts only when the previous tile's epilogue has drained (tiles serialise;
                    // the operand ring keeps streaming meanwhile)
                    if (it > 0u) { const uint32_t j = it - 1u; tc::mbar_wait(&sh->tmem_empty[j % (uint32_t)G], (j / (uint32_t)G) & 1u); tc::fence_after_sync(); }
                    constexpr int NB = 512 / Cfg::TMEM_STRIDE;
                    const int kbpf = p.cin[s] / BLOCK_K;
                    const int parts = (NB - 1) < kbpf ? (NB - 1) : kbpf;
                    uint32_t used = 0u;
                    int kf = 0, kbi = 0;
                    for (int kb = 0; kb < nkb; ++kb) {
                        const int acc = (kf + 1 < p.K_frames) ? 0 : 1 + (kbi * parts) / kbpf;
                        const uint32_t d_acc = tmem_base + (uint32_t)acc * Cfg::TMEM_STRIDE;
                        tc::mbar_wait(&sh->full[stage], phase);
                        tc::fence_after_sync();
                        const uint32_t a_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
                        const uint64_t da = tc::make_smem_desc_sw128(a_addr);
                        const uint64_t db = tc::make_smem_desc_sw128(a_addr + A_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            tc::umma_bf16(d_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)(k != 0 || ((used >> acc) & 1u)));
                        used |= 1u << acc;
                        tc::umma_commit(&sh->empty[stage]);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                        if (++kbi == kbpf) { kbi = 0; ++kf; }
                    }
                    tc::umma_commit(&sh->tmem_full[buf]);
                    continue;
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    tc::mbar_wait(&sh->full[stage], phase);
                    tc::fence_after_sync();
                    const uint32_t a_addr = tc::smem_u32(ring + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = tc::make_smem_desc_sw128(a_addr);
                    const uint64_t db = tc::make_smem_desc_sw128(a_addr + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)      // +32 bytes along K inside the swizzle atom
                        tc::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                    tc::umma_commit(&sh->empty[stage]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                }
                tc::umma_commit(&sh->tmem_full[buf]);
                if (p.stamps && blockIdx.x == 0 && it < 250u) p.stamps[it * 16 + 2] = clock64();
            }
        }
    } else if (warp >= kEpiWarp0) {
        // =========================== epilogue: 2 groups x 128 threads ===========================
        const int wgrp = (warp - kEpiWarp0) >> 2;             // warpgroup index
        const int grp = wgrp % G;                             // the TMEM buffer this warpgroup drains (SPLIT warpgroups share one)
        const int half = wgrp / G;                            // which share of the class chunks it takes
        (void)half;
        const int et = threadIdx.x - kEpiWarp0 * 32 - wgrp * kEpiThreads;
        const int q = warp & 3;                               // TMEM lane quarter this warp may read
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        EpiGroupShared* gs = &sh->grp[grp];
        uint64_t* L = slist + (size_t)grp * Cfg::LIST_BUFS * kListCap;
        int sum_it = 0;
        (void)gs; (void)L; (void)sum_it; (void)et;
        uint32_t it = (uint32_t)grp;
        bool ws_ok = false;                                   // the workspace's histograms started this call zeroed (layout marker intact)
        if constexpr (EPI == EPI_FILTER) ws_ok = p.tile_counter[2] == p.ws_magic;
        (void)ws_ok;
        uint32_t spec_tb = 0u;                                // EPI_SPEC: the call's score threshold (float bits)
        (void)spec_tb;
        bool spec_ws_ok = false;
        if constexpr (EPI == EPI_SPEC) spec_ws_ok = p.tile_counter[2] == p.ws_magic;
        (void)spec_ws_ok;
        uint32_t spec_pend_base = 0u, spec_pend_n = 0u, spec_par = 0u; int spec_pend_f = 0;
        auto spec_flush = [&]() {                             // copy the previous tile's staged keys to their reserved range
            if constexpr (EPI == EPI_SPEC) {
                const uint32_t n = spec_pend_n;
                if (n) {
                    const uint32_t base = __shfl_sync(0xffffffffu, spec_pend_base, 0);
                    const uint64_t* src_s = slist + (size_t)((warp - kEpiWarp0) * 2 + (int)((spec_par ^ 1u) & 1u)) * kSpecStage;
                    uint64_t* dst = p.spec_lists + (size_t)spec_pend_f * kSpecCap;
                    VD_DEV_CHECK(spec_pend_f >= 0 && spec_pend_f < p.frames && n <= (uint32_t)kSpecStage);
                    for (uint32_t j = (uint32_t)lane; j < n; j += 32u) if (base + j < (uint32_t)kSpecCap) dst[base + j] = src_s[j];
                    spec_pend_n = 0u;
                }
            }
        };
        (void)spec_pend_base; (void)spec_pend_f; (void)spec_par;
        for (;; it += G) {
            const uint32_t slot = it % kSchedSlots;
            tc::mbar_wait(&sh->sched_full[slot], (it / kSchedSlots) & 1u);
            const int tile = sh->sched_tile[slot];
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->sched_empty[slot]);
            if (tile < 0) { spec_flush(); break; }
            int s, f, pblk; tile_coords(p, tile, s, f, pblk);
            const uint32_t buf = (uint32_t)grp;
            const int HW = p.g.HW[s], Wd = p.g.W[s];
            const int cell = pblk * BLOCK_M + q * 32 + lane;
            const bool inb = cell < HW;
            const float gx = (float)(cell % Wd), gy = (float)(cell / Wd);
            const float* bias = sbias + s * NPAD;
            const bool stamp = (EPI == EPI_FILTER || EPI == EPI_SPEC) && p.stamps && blockIdx.x == 0 && et == 0 && it < 250u;
            if (stamp) p.stamps[it * 16 + 3] = clock64();
            // global reads of the selection issued before the wait for the accumulator (their latency hides behind it)
            uint32_t pf_c0 = 0u, pf_c1 = 0u, pf_hint = 0u;
            if constexpr (EPI == EPI_SPEC) {                     // this frame slot's threshold
                const uint32_t floor_b = p.valid_thresh > 0.0f ? __float_as_uint(p.valid_thresh) : 0u;
                const uint32_t hint = spec_ws_ok ? __ldcg(p.spec_tau + f) : 0u;
                spec_tb = hint > floor_b ? hint : floor_b;
                if (spec_tb > 0x3f800001u) spec_tb = 0x3f800001u;
                // Nothing is emitted (threshold above every score) on a foreign workspace: the NMS kernel fails every frame
                // anyway, and with tau = the valid floor this was 8 ms of atomics for 64 dense frames.  (Also reading the frame's
                // running count here, to stop emitting once its list has overflowed, costs 2 us per step: the word is under
                // atomic traffic from every CTA.)
                if (!spec_ws_ok) spec_tb = 0x3f800001u;
            }
            if constexpr (EPI == EPI_FILTER) {
                if (ws_ok && et < 32) { const uint32_t* ch = p.coarse + (size_t)f * 64; pf_c0 = __ldcg(ch + 63 - 2 * lane); pf_c1 = __ldcg(ch + 62 - 2 * lane); }
                pf_hint = __ldg(p.hint_hi + f);
            }
            tc::mbar_wait(&sh->tmem_full[buf], (it / (uint32_t)G) & 1u);
            tc::fence_after_sync();
            if constexpr (EPI == EPI_SPEC) spec_flush();
            if (stamp) p.stamps[it * 16 + 4] = clock64();
            const uint32_t tbase = tmem_base + buf * Cfg::TMEM_STRIDE + lane_addr;
            if (p.split) {
                // fp32-parity modes: fold the partial accumulators into this tile's buffer, fp32 round-to-nearest adds (K-ranges of
                // p0 w0 first, the low-order products last), then run the ordinary epilogue on the sum.  A thread touches only its own
                // TMEM lane; with SPLIT warpgroups on one buffer each folds every SPLIT-th 16-column chunk, then they meet.
                constexpr int NB = 512 / Cfg::TMEM_STRIDE;
                const int kbpf = p.cin[s] / BLOCK_K;
                const int parts = (NB - 1) < kbpf ? (NB - 1) : kbpf;
                const uint32_t t0 = tmem_base + lane_addr;
#pragma unroll 1
                for (int c0 = 0; c0 < NPAD; c0 += 16) {
                    if (Cfg::SPLIT > 1 && ((c0 >> 4) % Cfg::SPLIT) != half) continue;
                    uint32_t acc[16], r[16];
                    tc::tmem_ld16(t0 + (uint32_t)(Cfg::TMEM_STRIDE + c0), acc); tc::tmem_ld_wait();
                    for (int j = 2; j <= parts; ++j) {
                        tc::tmem_ld16(t0 + (uint32_t)(j * Cfg::TMEM_STRIDE + c0), r); tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(__fadd_rn(__uint_as_float(acc[i]), __uint_as_float(r[i])));
                    }
                    tc::tmem_ld16(t0 + (uint32_t)c0, r); tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(__fadd_rn(__uint_as_float(acc[i]), __uint_as_float(r[i])));
                    tc::tmem_st16(tbase + (uint32_t)c0, acc);
                }
                tc::tmem_st_wait();
                if constexpr (Cfg::SPLIT > 1) {
                    if (grp == 0) asm volatile("bar.sync 5, %0;" ::"n"(Cfg::SPLIT * kEpiThreads) : "memory");
                    else asm volatile("bar.sync 6, %0;" ::"n"(Cfg::SPLIT * kEpiThreads) : "memory");
                }
            }

            if constexpr (EPI == EPI_FILTER || EPI == EPI_SPEC) {
                if (p.dbg == 1) {        // debug (VD_DEBUG_SKIP_EPILOGUE=1): mainloop only, accumulators dropped
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
                    continue;
                }
            }
            if constexpr (EPI == EPI_PRED) {
                float* out = p.pred[s] + (size_t)f * p.pred_frame_stride;
#pragma unroll 1
                for (int n0 = 0; n0 < NPAD; n0 += 16) {
                    uint32_t r[16];
                    tc::tmem_ld16(tbase + n0, r); tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        int n = n0 + i;
                        if (inb && n < p.n_valid) out[(size_t)n * HW + cell] = __uint_as_float(r[i]) + bias[n];
                    }
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
                continue;
            }

            // ---- per-anchor box + objectness (yolo3.py:172-177)
            float conf[3];
            uint32_t rb[3][5];
#pragma unroll
            for (int a = 0; a < 3; ++a) tc::tmem_ld<5>(tbase + a * P, rb[a]);
            tc::tmem_ld_wait();
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const uint32_t* r = rb[a];
                float tx = __uint_as_float(r[0]) + bias[a * P + 0], ty = __uint_as_float(r[1]) + bias[a * P + 1];
                float tw = __uint_as_float(r[2]) + bias[a * P + 2], th = __uint_as_float(r[3]) + bias[a * P + 3];
                float to = __uint_as_float(r[4]) + bias[a * P + 4];
                conf[a] = vd_sigmoid(to);
                Box4 bx;
                if constexpr (EPI == EPI_FILTER || EPI == EPI_SPEC) {
                    // the raw (tx,ty,tw,th) are stored; only the <= topk boxes that reach the NMS kernel are decoded there
                    // (same vd_decode_box, same bits) instead of all 3*HW of them here
                    if (!inb) conf[a] = __uint_as_float(0x7fc00000u);   // NaN: no score of a padding pixel passes `> valid_thresh`
                    VD_DEV_CHECK(f >= 0 && f < p.frames && s >= 0 && s < p.g.num_scales && (!inb || p.g.anc_base[s] + cell * 3 + a < p.g.anc_base[p.g.num_scales]));
                    // EPI_FILTER stores every anchor's record; EPI_SPEC only those of the (pixel, anchor) pairs that emit a candidate (below):
                    // ~600 of 10 647 records per frame instead of all (the full store cost 1.05 us of a 29.6 us step)
                    if (EPI == EPI_FILTER && inb && half == 0 && p.pass == 0 && p.dbg != 8) p.boxes[(size_t)f * p.g.anc_base[p.g.num_scales] + p.g.anc_base[s] + cell * 3 + a] = make_float4(tx, ty, tw, th);
                    (void)bx; (void)gx; (void)gy;
                } else {   // EPI_DET: class rows of this (cell, anchor)
                    bx = vd_decode_box(tx, ty, tw, th, gx, gy, p.g.stride[s], p.g.anchors[s][2 * a], p.g.anchors[s][2 * a + 1]);
                    const size_t rows_scale = (size_t)HW * 3;
                    float* drow = p.det + ((size_t)f * p.det_rows_total + p.g.row_base[s] + (size_t)cell * 3 + a) * 6;
                    // 8-column windows; the last one may read up to 7 columns past 3*P, which stays inside
                    // this accumulator buffer's TMEM_STRIDE columns (checked at compile time)
                    static_assert(2 * P + 5 + (C + 7) / 8 * 8 <= Cfg::TMEM_STRIDE, "class window leaves the TMEM buffer");
#pragma unroll 1
                    for (int c0 = 0; c0 < C; c0 += 8) {
                        uint32_t rc[8];
                        tc::tmem_ld8(tbase + a * P + 5 + c0, rc); tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int c = c0 + i;
                            if (c < C && c < p.c_valid && inb) {
                                float sc = vd_score(__uint_as_float(rc[i]) + bias[a * P + 5 + c], conf[a]);
                                float2* o = reinterpret_cast<float2*>(drow + (size_t)(p.c_off + c) * rows_scale * 6);
                                o[0] = make_float2(__fadd_rn(__fmul_rn(sc, 0.0f), (float)(p.c_off + c)), sc);
                                o[1] = make_float2(bx.x1, bx.y1);
                                o[2] = make_float2(bx.x2, bx.y2);
                            }
                        }
                    }
                }
            }
            if constexpr (EPI == EPI_DET) {
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
                continue;
            }

            if constexpr (EPI == EPI_SPEC) {
                // ---- speculative frame-level threshold: every candidate whose score can reach the call's threshold tau
                // (max of the valid floor and the hint left by the previous call) is scored and appended to its FRAME's
                // list; nothing is selected per tile, no barrier, no shared-memory list.  The NMS kernel verifies per frame
                // that the list did not overflow and holds >= k candidates at or above tau (then it contains the frame's
                // exact top-k); frames that fail are redone by the exact EPI_FILTER path.  Same conservative logit
                // prefilter as there: x >= logit(tau/conf) - margin  <=  score >= tau.
                constexpr int CH = Cfg::CH, CPA = Cfg::CPA, CH4 = Cfg::CH4, REM = C - (CPA - 1) * CH;
                if (stamp) p.stamps[it * 16 + 8] = clock64();          // box part done
                if (p.dbg == 7) { tc::fence_before_sync(); __syncwarp(); if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]); continue; }   // debug: box part only
                const float vth = p.dbg == 9 ? 2.0f : p.valid_thresh;                 // debug 9: nothing is ever emitted
                const float* cbias = scbias + s * (3 * CPA * CH4);
                float ell[3];
                {
                    const float t = __fmul_rn(__uint_as_float(spec_tb), 0.999969482421875f);   // 1 - 2^-15
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float r = __fmul_rn(t, vd_rcp(conf[a]));
                        const float l = __fmul_rn(__fsub_rn(vd_lg2(r), vd_lg2(__fsub_rn(1.0f, r))), 0.6931471805599453f);
                        float e = (r < 1.0f) ? l : __uint_as_float(0x7f800000u);
                        if (spec_tb == 0u) e = __uint_as_float(0xff800000u);
                        if (!(conf[a] == conf[a])) e = __uint_as_float(0x7f800000u);
                        if (spec_tb >= 0x3f800001u) e = __uint_as_float(0x7f800000u);      // "emit nothing" (see above): conf may exceed tau/1 only by rounding
                        ell[a] = e;
                    }
                }
                const uint32_t HW3 = (uint32_t)HW * 3u;
                const uint32_t row0 = (uint32_t)p.g.row_base[s] + (uint32_t)cell * 3u + (uint32_t)p.c_off * HW3;       // + c*HW3 + a
                const int cval = p.c_valid;
                uint64_t* fl = p.spec_lists + (size_t)f * kSpecCap;
                uint32_t* fc = p.spec_cnt + f;
                // Passers are staged per WARP in shared memory (no group barrier anywhere in this epilogue).  The warp's
                // range of the frame's list is reserved with one global atomic whose result is only consumed when the
                // NEXT tile's accumulator is ready, so its round trip hides behind that wait (deferred flush).
                const int wq = (warp - kEpiWarp0);                                  // epilogue warp index inside the CTA
                uint64_t* stg = slist + (size_t)(wq * 2 + (int)(spec_par & 1u)) * kSpecStage;
                uint32_t* wc = &sh->spec_wcnt[wq];
                uint32_t r[2][CH];
                auto issue_a = [&](const int a, const int cc, uint32_t* dst) {
                    const uint32_t col = tbase + (uint32_t)(a * P + 5 + cc * CH);
                    if (REM != CH && cc == CPA - 1) tc::tmem_ld<REM>(col, dst); else tc::tmem_ld<CH>(col, dst);
                };
                uint32_t emit_mask = 0u;                              // anchors of this pixel that emitted a candidate from this tile
                auto emit_chunk = [&](const float* bv, const int n, const int a, const int cc, const float la, const float ca) {
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        if (i < n && bv[i] >= la && cc * CH + i < cval) {
                            const float sc = vd_score(bv[i], ca);
                            if (sc > vth) {
                                emit_mask |= 1u << a;
                                const uint32_t kh = __float_as_uint(sc) | 0x80000000u;
                                const uint32_t row = row0 + (uint32_t)(cc * CH + i) * HW3 + (uint32_t)a;
                                VD_DEV_CHECK(row < (uint32_t)p.g.row_base[p.g.num_scales] && inb);
                                const uint64_t key = ((uint64_t)kh << 32) | (uint32_t)~row;
                                const uint32_t sp = atomicAdd(wc, 1u);
                                if (sp < (uint32_t)kSpecStage) stg[sp] = key;
                                else { const uint32_t pos = atomicAdd(fc, 1u); if (pos < (uint32_t)kSpecCap) fl[pos] = key; }   // staging full
                            }
                        }
                    }
                };
                if constexpr (Cfg::SPLIT > 1) {
                    // flat walk over the 3*CPA class chunks, this warpgroup takes chunks half, half+SPLIT, ...; the next
                    // chunk's TMEM read is in flight while the current one is compared
                    constexpr int NCH = 3 * CPA;
                    auto issue_j = [&](const int j, uint32_t* dst) {
                        const int a = j / CPA, cc = j - a * CPA;
                        const uint32_t col = tbase + (uint32_t)(a * P + 5 + cc * CH);
                        if (REM != CH && cc == CPA - 1) tc::tmem_ld<REM>(col, dst); else tc::tmem_ld<CH>(col, dst);
                    };
                    if (half < NCH) issue_j(half, r[0]);
                    int par = 0;
#pragma unroll 1
                    for (int j = half; j < NCH; j += Cfg::SPLIT, par ^= 1) {
                        const int a = j / CPA, cc = j - a * CPA;
                        const int n = (REM != CH && cc == CPA - 1) ? REM : CH;
                        const float la = (a == 0) ? ell[0] : ((a == 1) ? ell[1] : ell[2]);
                        const float ca = (a == 0) ? conf[0] : ((a == 1) ? conf[1] : conf[2]);
                        float bv[CH4];
#pragma unroll
                        for (int i = 0; i < CH4; i += 4)
                            *reinterpret_cast<float4*>(bv + i) = *reinterpret_cast<const float4*>(cbias + j * CH4 + i);
                        tc::tmem_ld_wait();
                        bool any = false;
                        if (par == 0) {
                            if (j + Cfg::SPLIT < NCH) issue_j(j + Cfg::SPLIT, r[1]);
#pragma unroll
                            for (int i = 0; i < CH; ++i) if (i < n) { bv[i] = __fadd_rn(__uint_as_float(r[0][i]), bv[i]); any |= bv[i] >= la; }
                        } else {
                            if (j + Cfg::SPLIT < NCH) issue_j(j + Cfg::SPLIT, r[0]);
#pragma unroll
                            for (int i = 0; i < CH; ++i) if (i < n) { bv[i] = __fadd_rn(__uint_as_float(r[1][i]), bv[i]); any |= bv[i] >= la; }
                        }
                        if (any) emit_chunk(bv, n, a, cc, la, ca);
                    }
                } else {
                issue_a(0, 0, r[0]);
#pragma unroll 1
                for (int a = 0; a < 3; ++a) {
                    const float la = (a == 0) ? ell[0] : ((a == 1) ? ell[1] : ell[2]);
                    const float ca = (a == 0) ? conf[0] : ((a == 1) ? conf[1] : conf[2]);
#pragma unroll
                    for (int cc = 0; cc < CPA; ++cc) {
                        const int n = (REM != CH && cc == CPA - 1) ? REM : CH;
                        float bv[CH4];
#pragma unroll
                        for (int i = 0; i < CH4; i += 4)
                            *reinterpret_cast<float4*>(bv + i) = *reinterpret_cast<const float4*>(cbias + (a * CPA + cc) * CH4 + i);
                        if ((CPA & 1) && cc == 0 && a > 0) issue_a(a, 0, r[0]);
                        tc::tmem_ld_wait();
                        if (cc + 1 < CPA) issue_a(a, cc + 1, r[(cc + 1) & 1]);
                        else if (!(CPA & 1) && a + 1 < 3) issue_a(a + 1, 0, r[0]);
                        bool any = false;
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            if (i < n) { bv[i] = __fadd_rn(__uint_as_float(r[cc & 1][i]), bv[i]); any |= bv[i] >= la; }
                        }
                        if (any) emit_chunk(bv, n, a, cc, la, ca);   // rare: ~0.3 % of the candidates pass
                    }
                }
                }
                if (stamp) p.stamps[it * 16 + 10] = clock64();         // class loop done
                // raw box records of the emitting (pixel, anchor) pairs only: re-read from the accumulator (same bits as the box part);
                // tcgen05.ld is warp-collective, so a warp reads an anchor's columns when any of its lanes needs them
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const bool mine = ((emit_mask >> a) & 1u) != 0u;
                    if (__any_sync(0xffffffffu, mine) && p.dbg != 8) {
                        uint32_t r4[4];
                        tc::tmem_ld<4>(tbase + (uint32_t)(a * P), r4); tc::tmem_ld_wait();
                        if (mine) p.boxes[(size_t)f * p.g.anc_base[p.g.num_scales] + p.g.anc_base[s] + cell * 3 + a] =
                            make_float4(__uint_as_float(r4[0]) + bias[a * P + 0], __uint_as_float(r4[1]) + bias[a * P + 1],
                                        __uint_as_float(r4[2]) + bias[a * P + 2], __uint_as_float(r4[3]) + bias[a * P + 3]);
                    }
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
                // reserve this tile's range now, copy it out when the next accumulator arrives
                uint32_t wn = *wc; wn = wn < (uint32_t)kSpecStage ? wn : (uint32_t)kSpecStage;
                __syncwarp();
                spec_pend_base = 0u;
                if (lane == 0) { *wc = 0u; if (wn) spec_pend_base = atomicAdd(fc, wn); }
                spec_pend_n = wn; spec_pend_f = f; spec_par ^= 1u;
                __syncwarp();
                if (stamp) p.stamps[it * 16 + 7] = clock64();
                continue;
            }

            if constexpr (EPI == EPI_FILTER) {
                // ---- exact selection of the tile's top-k WITHOUT scoring every candidate.
                // score = sigmoid(x)*conf is monotone in the class logit x, so "score >= t" <=> "x >= logit(t/conf)":
                // one threshold per (pixel, anchor) turns each probe of the pivot search into a compare on the raw
                // accumulator.  The threshold is made conservative by 2^-15 relative in score space (>> the ~5e-6
                // error of the approx ex2/rcp/lg2 chain), so the prefilter set P(t) contains every candidate whose
                // COMPUTED score is >= t.  Only members of P (3-5 per thread) are scored; the tile's list is
                // accepted when  k <= #{computed score >= t}  and  |P| <= cap  => list is a superset of the tile's
                // exact top-k.  Heavily tied scores (fast probes cannot separate) fall back to an exact bisection
                // on the 64-bit (score,row) keys, which scores every candidate per probe (rare).
                //   sweep 1  count |P| (compare only)            -> block sum -> accept / move the threshold
                //   sweep 2  stage P's (logit, local code) at exact list positions (predicated stores)
                //   score    the staged entries, 128 threads striding the list (balanced), keys kept in registers
                //   flush    keys -> global tile list (true row restored) + per-frame score histogram
                if (p.dbg == 2) { tc::fence_before_sync(); __syncwarp(); if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]); continue; }
                const uint32_t k = (uint32_t)p.k, cap = (uint32_t)p.cap;
                const uint32_t cellofs = (uint32_t)(q * 32 + lane);
                const uint32_t lcode = cellofs << 2;                // local code of (class c, anchor a) = lcode | (c << 9) | a; key low word = ~code
                const int slot = 2 * s + ((pblk + 1) * BLOCK_M > HW ? 1 : 0);
                const float vth = p.valid_thresh;
                // Running lower bound of the frame's k-th best score: the coarse histogram counts candidates that tiles of this
                // frame have ALREADY emitted, so the highest coarse bin whose suffix count reaches k is a score that at least
                // k candidates of the frame attain -- nothing below it can be in the frame's top-k.  It becomes this tile's
                // floor ("emit everything above the floor if it fits" needs no search); any stale view of the counts is safe.
                uint32_t floor_b = vth > 0.0f ? __float_as_uint(vth) : 0u;         // lowest threshold (float bits); 0: everything
                if (et < 64) gs->chist[et] = 0u;
                if (ws_ok) {                                                       // (uniform) the coarse histogram started this call zeroed
                    if (et < 32) {                                                 // one warp reads it: the whole group must agree on the floor
                        const uint32_t c0 = pf_c0, c1 = pf_c1;
                        uint32_t incl = c0 + c1;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                        const uint32_t excl = incl - (c0 + c1);
                        const unsigned hit = __ballot_sync(0xffffffffu, incl >= (uint32_t)p.k);
                        uint32_t bound = 0u;
                        if (hit) {
                            const int src = __ffs(hit) - 1;
                            const uint32_t mybin = (excl + c0 >= (uint32_t)p.k) ? (uint32_t)(63 - 2 * lane) : (uint32_t)(62 - 2 * lane);
                            const uint32_t bin = __shfl_sync(0xffffffffu, mybin, src);
                            bound = hist_edge(bin << 6) & 0x7fffffffu;             // score bits of the coarse bin's lower edge
                        }
                        if (lane == 0) gs->bound = bound;
                    }
                    epi_bar(grp);
                    const uint32_t bound = gs->bound;
                    if (bound > floor_b) floor_b = bound;
                }
                constexpr uint32_t kTop = 0x3f800001u;              // > 1.0: nothing scores above it
                constexpr int CH = Cfg::CH, CPA = Cfg::CPA, CH4 = Cfg::CH4, REM = C - (CPA - 1) * CH, NCHUNK = 3 * CPA;
                const float* cbias = scbias + s * (3 * CPA * CH4);
                float* cf = sconf + grp * (3 * kEpiThreads);
#pragma unroll
                for (int a = 0; a < 3; ++a) cf[cellofs * 3 + a] = conf[a];   // visible to the group after the first block sum
                float ell[3];
                auto set_fast = [&](uint32_t tb) {
                    const float t = __fmul_rn(__uint_as_float(tb), 0.999969482421875f);   // 1 - 2^-15
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float r = __fmul_rn(t, vd_rcp(conf[a]));
                        const float l = __fmul_rn(__fsub_rn(vd_lg2(r), vd_lg2(__fsub_rn(1.0f, r))), 0.6931471805599453f);
                        float e = (r < 1.0f) ? l : __uint_as_float(0x7f800000u);          // r >= 1 or NaN: nothing passes
                        if (tb == 0u) e = __uint_as_float(0xff800000u);                   // no positive threshold: everything passes
                        if (!(conf[a] == conf[a])) e = __uint_as_float(0x7f800000u);      // padding pixel
                        ell[a] = e;
                    }
                };
                auto issue = [&](const int j, uint32_t* r) {
                    const int a = j / CPA, cc = j - a * CPA;
                    const uint32_t col = tbase + (uint32_t)(a * P + 5 + cc * CH);
                    if (REM != CH && cc == CPA - 1) tc::tmem_ld<REM>(col, r); else tc::tmem_ld<CH>(col, r);
                };
                // fast sweep: ONE pass over the class logits in TMEM (next chunk's load in flight while this one is compared).
                // Per chunk a thread counts its passers, the warp scans the counts and reserves a range of the group's list
                // with one shared-memory atomic, and the passers' (logit, code) entries are stored from the registers that
                // still hold them.  The list is speculative: the caller accepts it iff the final cursor is inside the band.
                const uint32_t L_s = tc::smem_u32(L);
                auto sweep_fast = [&](uint32_t* cursor) {
                    uint32_t r[2][CH];
                    issue(0, r[0]);
#pragma unroll
                    for (int j = 0; j < NCHUNK; ++j) {
                        const int a = j / CPA, cc = j - a * CPA;
                        const int n = (REM != CH && cc == CPA - 1) ? REM : CH;
                        float bv[CH4];
#pragma unroll
                        for (int i = 0; i < CH4; i += 4)
                            *reinterpret_cast<float4*>(bv + i) = *reinterpret_cast<const float4*>(cbias + (a * CPA + cc) * CH4 + i);
                        tc::tmem_ld_wait();
                        if (j + 1 < NCHUNK) issue(j + 1, r[(j + 1) & 1]);
                        const float la = ell[a];
                        uint32_t c0 = 0u, c1 = 0u;
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            if (i < n) {
                                bv[i] = __fadd_rn(__uint_as_float(r[j & 1][i]), bv[i]);      // the logit, kept for the stores below
                                if (i & 1) c1 += (bv[i] >= la) ? 1u : 0u; else c0 += (bv[i] >= la) ? 1u : 0u;
                            }
                        }
                        const uint32_t mine = c0 + c1;
                        uint32_t incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                        if (tot) {                                   // warp-uniform
                            uint32_t base = 0u;
                            if (lane == 31) base = atomicAdd(cursor, tot);
                            base = __shfl_sync(0xffffffffu, base, 31);
                            uint32_t pos = base + incl - mine;
                            if (mine) {
#pragma unroll
                                for (int i = 0; i < CH; ++i) {
                                    if (i < n) {
                                        const bool pass = (bv[i] >= la) & (pos < (uint32_t)kListCap);
                                        const uint32_t code = lcode | (((uint32_t)(cc * CH + i)) << 9) | (uint32_t)a;
                                        if (pass) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(L_s + pos * 8u), "r"(code), "r"(__float_as_uint(bv[i])) : "memory");
                                        pos += (bv[i] >= la) ? 1u : 0u;
                                    }
                                }
                            }
                        }
                    }
                };
                // exact sweep (rare): every candidate is scored; counts / writes the keys >= piv
                auto sweep_exact = [&](auto emit_tag, uint64_t* wp, uint64_t* wend, const uint64_t piv) -> uint32_t {
                    constexpr bool EMIT = decltype(emit_tag)::value;
                    static_assert(2 * P + 5 + (C + 7) / 8 * 8 <= Cfg::TMEM_STRIDE, "class window leaves the TMEM buffer");
                    uint32_t cnt = 0u;
#pragma unroll 1
                    for (int a = 0; a < 3; ++a) {
                        const float ca = cf[cellofs * 3 + a];
#pragma unroll 1
                        for (int c0 = 0; c0 < C; c0 += 8) {
                            uint32_t rc[8];
                            tc::tmem_ld8(tbase + (uint32_t)(a * P + 5 + c0), rc); tc::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int c = c0 + i;
                                if (c < C) {
                                    const float sc = vd_score(__fadd_rn(__uint_as_float(rc[i]), bias[a * P + 5 + c]), ca);
                                    const uint32_t kh = (sc > vth && c < p.c_valid) ? (__float_as_uint(sc) | 0x80000000u) : 0u;
                                    const uint64_t key = ((uint64_t)kh << 32) | (uint32_t)~(lcode | (((uint32_t)c) << 9) | (uint32_t)a);
                                    if (kh && key >= piv) {
                                        ++cnt;
                                        if constexpr (EMIT) { if (wp < wend) *wp++ = key; }
                                    }
                                }
                            }
                        }
                    }
                    return cnt;
                };
                // position of this thread's entries in the group's list: warp scan + one cursor bump per warp
                auto list_base = [&](const uint32_t mine) -> uint32_t {
                    uint32_t incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                    uint32_t base = 0;
                    if (lane == 31 && tot) base = atomicAdd(&gs->cursor, tot);
                    base = __shfl_sync(0xffffffffu, base, 31);
                    return base + incl - mine;
                };
                auto release_tmem = [&]() {
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
                };
                // global tile list + histogram
                const size_t li = ((size_t)f * p.n_pass + p.pass) * p.tiles_per_frame + p.tif_base[s] + pblk;
                uint64_t* gl = p.lists + li * kListCap;
                const uint32_t HW3 = (uint32_t)HW * 3u;
                const uint32_t rbase = (uint32_t)p.g.row_base[s] + (uint32_t)(pblk * BLOCK_M) * 3u + (uint32_t)p.c_off * HW3;
                auto flush_one = [&](const uint32_t j, const uint64_t v) {     // v: (key_hi, ~local code) or 0
                    uint64_t o = 0ull;
                    if (v != 0ull) {
                        const uint32_t local = ~(uint32_t)v;
                        const uint32_t row = rbase + (local >> 9) * HW3 + ((local >> 2) & 127u) * 3u + (local & 3u);
                        o = (v & 0xffffffff00000000ull) | (uint32_t)(~row);
                        const uint32_t hb = hist_bin((uint32_t)(v >> 32));
                        VD_DEV_CHECK(hb < (uint32_t)kHistBins && row < (uint32_t)p.g.row_base[p.g.num_scales] && j < (uint32_t)kListCap);
                        if (p.dbg != 5) { atomicAdd(&p.hist[(size_t)f * kHistBins + hb], 1u); atomicAdd(&gs->chist[hb >> 6], 1u); }
                    }
                    if (p.dbg != 6) gl[j] = o;
                };

                if (stamp) p.stamps[it * 16 + 8] = clock64();          // box decode done
                uint32_t guess = (uint32_t)gs->guess[slot], band_w = (uint32_t)gs->band[slot];
                if (band_w == 0u || band_w > (1u << 26)) band_w = 1u << 18;
                uint32_t piv = guess;
                if (piv < floor_b) piv = floor_b;
                if (piv > kTop) piv = kTop;
                uint32_t lo = floor_b, hi = kTop, step = band_w;
                bool have_lo = false, have_hi = false, exact = false;
                uint32_t list_n = 0, list_hi = 0, t_acc = 0, my_cnt = 0;
                uint64_t piv64 = 0ull, lo64 = 0ull, hi64 = 0ull;
#pragma unroll 1
                for (int g = 0; g < 200; ++g) {
                    if (!exact) {
                        set_fast(piv);
                        if (stamp) p.stamps[it * 16 + 9] = clock64();      // thresholds set
                        uint32_t* pc = &gs->pcur[g % 3];                   // this probe's list cursor (zero on entry)
                        if (et == 0) gs->pcur[(g + 1) % 3] = 0u;           // next probe's: last read two probes ago
                        sweep_fast(pc);
                        if (stamp) p.stamps[it * 16 + 10] = clock64();     // sweep done
                        epi_bar(grp);                                      // every reservation and store of the group has landed
                        const uint32_t t = *pc;
                        if (stamp) p.stamps[it * 16 + 5] = clock64();
                        if (p.dbg == 3) { release_tmem(); break; }
                        if (t <= cap && (t >= k || piv == floor_b)) {
                            // ---- the staged list is P(piv): score it, check the exact count
                            const uint32_t ph = (piv == floor_b) ? 1u : (piv | 0x80000000u);
                            // score; split the list at the frame's hint: keys >= hint_hi go to the front, so the NMS kernel
                            // reads only the fronts when its pivot is at or above the hint (the usual case)
                            constexpr int KPT = kListCap / kEpiThreads;
                            const uint32_t thk = pf_hint;
                            uint64_t key[KPT];
                            uint32_t off[KPT];                      // bit 31: front part; low bits: offset inside this warp's share of the part
                            uint32_t ne = 0u, nh = 0u, wh = 0u, wl = 0u;
#pragma unroll
                            for (int u = 0; u < KPT; ++u) {
                                const uint32_t j = (uint32_t)(u * kEpiThreads + et);
                                key[u] = 0ull; off[u] = 0u;
                                if ((uint32_t)(u * kEpiThreads) < t) {      // group-uniform
                                    const bool have = j < t;
                                    uint32_t kh = 0u;
                                    if (have) {
                                        const uint64_t e = L[j];
                                        const uint32_t code = (uint32_t)e;
                                        const float sc = vd_score(__uint_as_float((uint32_t)(e >> 32)), cf[((code >> 2) & 127u) * 3u + (code & 3u)]);
                                        kh = (sc > vth && (int)(code >> 9) < p.c_valid) ? (__float_as_uint(sc) | 0x80000000u) : 0u;
                                        key[u] = kh ? (((uint64_t)kh << 32) | (uint32_t)~code) : 0ull;
                                        ne += (kh >= ph) ? 1u : 0u;
                                    }
                                    const bool hi = have && kh != 0u && kh >= thk;
                                    const unsigned bh = __ballot_sync(0xffffffffu, hi), bl = __ballot_sync(0xffffffffu, have && !hi);
                                    const unsigned lt = (1u << lane) - 1u;
                                    off[u] = hi ? (0x80000000u | (wh + (uint32_t)__popc(bh & lt))) : (wl + (uint32_t)__popc(bl & lt));
                                    wh += (uint32_t)__popc(bh); wl += (uint32_t)__popc(bl);
                                    nh += hi ? 1u : 0u;
                                }
                            }
                            uint32_t wbase = 0u;
                            if (lane == 0) wbase = atomicAdd(&gs->cursor2, wh | (wl << 16));   // this warp's share of the front / back parts
                            wbase = __shfl_sync(0xffffffffu, wbase, 0);
                            if (stamp) p.stamps[it * 16 + 13] = clock64(); // scored
                            const uint32_t t2p = epi_sum(ne | (nh << 16), gs, grp, et, sum_it);
                            const uint32_t t2 = t2p & 0xffffu, th = t2p >> 16;
                            if (p.dbg == 4) { release_tmem(); break; }
                            if (t2 >= k || piv == floor_b) {
                                release_tmem();
                                if (stamp) p.stamps[it * 16 + 6] = clock64();
#pragma unroll
                                for (int u = 0; u < KPT; ++u) {
                                    const uint32_t j = (uint32_t)(u * kEpiThreads + et);
                                    if (j < t) {
                                        const uint32_t o = off[u] & 0x7fffffffu;
                                        flush_one((off[u] >> 31) ? (wbase & 0xffffu) + o : th + (wbase >> 16) + o, key[u]);
                                    }
                                }
                                list_n = t; t_acc = t2; list_hi = th;
                                if (stamp) p.stamps[it * 16 + 14] = clock64(); // flushed
                                break;
                            }
                            if (et == 0) gs->cursor2 = 0;
                            // margin ate the k-th candidates (tie cluster right at the threshold): exact search just below
                            if (et == 0) gs->cursor = 0;
                            exact = true;
                            hi64 = (uint64_t)(piv | 0x80000000u) << 32;                       // count(hi64) = t2 < k
                            const uint32_t lb = __float_as_uint(__fmul_rn(__uint_as_float(piv), 0.9998779296875f));   // 1 - 2^-13
                            lo64 = (lb <= floor_b) ? 1ull : ((uint64_t)(lb | 0x80000000u) << 32);
                            piv64 = lo64;
                            continue;
                        }
                        if (t > cap) { lo = piv; have_lo = true; } else { hi = piv; have_hi = true; }
                        if (have_lo && lo >= kTop) { have_hi = true; hi = kTop; }   // scores saturated at 1.0: exact keys decide
                        if (have_lo && have_hi) {
                            if (hi - lo <= 1024u) {                 // below the prefilter's resolution: exact keys decide
                                exact = true;
                                hi64 = (uint64_t)(hi | 0x80000000u) << 32;
                                const uint32_t lb = __float_as_uint(__fmul_rn(__uint_as_float(lo), 0.9998779296875f));
                                lo64 = (lo == floor_b || lb <= floor_b) ? 1ull : ((uint64_t)(lb | 0x80000000u) << 32);
                                piv64 = lo64;
                                continue;
                            }
                            if (hi - lo < band_w) band_w = (hi - lo) | 1024u;
                            piv = lo + ((hi - lo) >> 1);
                        } else if (have_lo) {                       // gallop up
                            const uint64_t nm = (uint64_t)piv + step;
                            piv = nm >= (uint64_t)kTop ? kTop : (uint32_t)nm;
                            step = step < (1u << 30) ? step << 1 : step;
                        } else {                                    // gallop down, never below the floor
                            piv = (piv > floor_b && piv - floor_b > step) ? piv - step : floor_b;
                            step = step < (1u << 30) ? step << 1 : step;
                        }
                    } else {
                        my_cnt = sweep_exact(std::false_type{}, nullptr, nullptr, piv64);
                        const uint32_t t = epi_sum(my_cnt, gs, grp, et, sum_it);
                        const bool last = (hi64 - lo64 <= 1ull) || g >= 198;
                        if ((t <= cap && (t >= k || piv64 == 1ull)) || last) {
                            const uint32_t pos = list_base(my_cnt);
                            const uint32_t pc = pos < (uint32_t)kListCap ? pos : (uint32_t)kListCap;
                            sweep_exact(std::true_type{}, L + pc, L + kListCap, piv64);
                            release_tmem();
                            epi_bar(grp);                           // list complete
                            list_n = t < (uint32_t)kListCap ? t : (uint32_t)kListCap; t_acc = t; list_hi = list_n;   // no split: everything counts as front
                            for (uint32_t j = et; j < list_n; j += kEpiThreads) flush_one(j, L[j]);
                            piv = (uint32_t)(piv64 >> 32) & 0x7fffffffu;
                            break;
                        }
                        if (t > cap) lo64 = piv64; else hi64 = piv64;
                        piv64 = lo64 + ((hi64 - lo64) >> 1);
                    }
                }
                // ---- retarget the warm start towards the band centre
                if (et == 0) {
                    gs->pcur[0] = gs->pcur[1] = gs->pcur[2] = 0u; gs->cursor = 0u; gs->cursor2 = 0u;   // next tile starts from empty lists
                    VD_DEV_CHECK(li < (size_t)p.frames * p.n_pass * p.tiles_per_frame && list_n <= (uint32_t)kListCap && list_hi <= list_n);
                    p.counts[li] = list_n; p.counts_hi[li] = list_hi;
                    if (piv > floor_b) {
                        const uint32_t q4 = (cap - k) / 4u, nudge = band_w >> 3;
                        uint32_t g2 = piv;
                        if (t_acc < k + q4 && piv > floor_b + nudge) g2 = piv - nudge;
                        else if (list_n > cap - q4) g2 = piv + nudge;
                        gs->guess[slot] = g2; gs->band[slot] = band_w;
                        if (p.hints) { p.hints[2 * slot] = g2; p.hints[2 * slot + 1] = band_w; }   // racy by design: any value is only a hint
                    }
                }
                epi_bar(grp);                                       // list buffer / cursor / conf table / guess slots are reused by the next tile
                if (et < 64) { const uint32_t cc = gs->chist[et]; if (cc) atomicAdd(&p.coarse[(size_t)f * 64 + et], cc); }
                if (stamp) p.stamps[it * 16 + 7] = clock64();
            }
        }
    }

    // ---------------- teardown
    __syncwarp();                // warps 0 / 1 ran single-lane role loops: reconverge before the aligned CTA barrier
    tc::fence_before_sync();
    __syncthreads();
    if ((EPI == EPI_FILTER || EPI == EPI_SPEC) && p.stamps && p.frame_list == nullptr && threadIdx.x == 0) p.stamps[4096 + blockIdx.x * 4 + 2] = gtime();
    if (warp == 1) { if (p.split) tc::tmem_dealloc<512>(tmem_base); else tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
    if (threadIdx.x == 0 && p.tile_counter != nullptr && p.tile_counter[2] == p.ws_magic) {
        // last CTA out re-arms the scheduler for the next launch on this workspace
        __threadfence();
        if (atomicAdd(&p.tile_counter[1], 1u) == gridDim.x - 1u) { p.tile_counter[0] = 0u; p.tile_counter[1] = 0u; }
    }
}

// ----------------------------------------------------------------------------------------------
// per-frame top-k + NMS on the tile lists (fused-path Source / Sink for nms_final_body)
// ----------------------------------------------------------------------------------------------
struct FusedSource {
    HeadGeom g; const float4* boxes;
    __device__ __forceinline__ void load(int f, uint32_t row, float, float4& bx, int& c, float& area) const {
        int s = 0;
#pragma unroll
        for (int i = 1; i < VD_MAX_SCALES; ++i) if (i < g.num_scales && (int)row >= g.row_base[i]) s = i;
        const int rs = (int)row - g.row_base[s];
        const int per = g.HW[s] * 3;
        c = rs / per;
        const int slot = rs - c * per;
        VD_DEV_CHECK((int)row < g.row_base[g.num_scales] && rs >= 0 && c < g.num_class && slot < per);
        const float4 t = boxes[(size_t)f * g.anc_base[g.num_scales] + g.anc_base[s] + slot];     // raw (tx,ty,tw,th) of the anchor
        const int cell = slot / 3, a = slot - cell * 3;
        const int gy = cell / g.W[s], gx = cell - gy * g.W[s];
        const Box4 b = vd_decode_box(t.x, t.y, t.z, t.w, (float)gx, (float)gy, g.stride[s], g.anchors[s][2 * a], g.anchors[s][2 * a + 1]);
        bx = make_float4(b.x1, b.y1, b.x2, b.y2);
        area = vd_box_area(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    }
};
// The sink of the fused path.  Output mirrors (VdHeadParams::n_mirrors): every ids / scores / bboxes element stored at address a
// is also stored at a + mirror[i] -- the same slot of the detection gather buffer of peer GPU i, mapped into this process over
// NVLink (CUDA IPC): the final detection gather of the multi-GPU job happens inside the NMS kernel, with no staging copy and
// no collective on the data path (2.4 KB per frame and peer).
struct FusedSink {
    float* ids; float* scores; float* bboxes; int32_t* keep; int post;
    int n_mirror; long long mirror[VD_MAX_MIRRORS];
    template <class T> __device__ __forceinline__ void put(T* p, const T v) const {
        *p = v;
        for (int i = 0; i < n_mirror; ++i) *reinterpret_cast<T*>(reinterpret_cast<char*>(p) + mirror[i]) = v;
    }
    __device__ __forceinline__ void emit(int f, int pos, uint32_t row, float score, float4 bx, int c) const {
        size_t o = (size_t)f * post + pos;
        VD_DEV_CHECK(pos >= 0 && pos < post && f >= 0);
        put(ids + o, __fadd_rn(__fmul_rn(score, 0.0f), (float)c));
        put(scores + o, score);
        put(reinterpret_cast<float4*>(bboxes) + o, bx);
        if (keep) keep[o] = (int32_t)row;
    }
    __device__ __forceinline__ void finish(int f, int kept) const {
        for (int pos = kept + threadIdx.x; pos < post; pos += blockDim.x) {
            size_t o = (size_t)f * post + pos;
            put(ids + o, -1.0f); put(scores + o, -1.0f);
            put(reinterpret_cast<float4*>(bboxes) + o, make_float4(-1.f, -1.f, -1.f, -1.f));
            if (keep) keep[o] = -1;
        }
    }
};

constexpr int kMaxFrameLists = 512;          // per-tile candidate lists of one frame the exact NMS kernel can walk (pixel blocks x class windows)

// Per-frame top-k + NMS straight from the tile lists: the frame pivot comes from the score histogram
// (suffix scan), the lists are streamed once and compacted, then sort + nms_tail.  No merge passes,
// no pivot search (a streaming bisection remains as the fallback for pathologically tied scores or a
// histogram that does not describe the lists).  Sized to share an SM with a head_kernel CTA of the
// NEXT batch (256 threads, <= 64 registers, ~41 KB shared): the step pipeline overlaps the two.
// The kernel leaves its frame's histogram zeroed for the next call (no memset node per call).
__global__ void __maxnreg__(56)
nms_final_hist_kernel(const uint64_t* __restrict__ lists, const uint32_t* __restrict__ counts, const uint32_t* __restrict__ counts_hi,
                      uint32_t* __restrict__ hint_hi, uint32_t* __restrict__ coarse, int n_lists,
                      uint32_t* __restrict__ hist, unsigned int* __restrict__ ctr, unsigned int ws_magic,
                      const uint32_t* __restrict__ frame_list, uint32_t* __restrict__ spec_state, uint32_t* __restrict__ spec_tau, NmsParams P, FusedSource src, FusedSink sink) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int k = P.k;
    // fallback of the speculative path: CTA i takes the i-th failed frame; every CTA joins the end-of-call bookkeeping
    const bool active = frame_list == nullptr || blockIdx.x < spec_state[2];
    const int f = frame_list ? (active ? (int)frame_list[blockIdx.x] : 0) : (int)blockIdx.x;
    if (active) {
    SelectScratch* scr = reinterpret_cast<SelectScratch*>(smem_raw);
    uint32_t* sx = reinterpret_cast<uint32_t*>(smem_raw + 64);          // [40] warp sums + results
    uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + 256);
    float4* sbox = reinterpret_cast<float4*>(skeys + kHistCap);
    const int KMAX = (P.max_out < k ? P.max_out : k) + 32 + 4 * (kNmsThreads / 32);   // kept list + padding
    float4* skbox = sbox + k;
    float* skarea = reinterpret_cast<float*>(skbox + KMAX);
    int* skcls = reinterpret_cast<int*>(skarea + KMAX);
    int* scls = skcls + KMAX;
    float* sarea = reinterpret_cast<float*>(scls + k);
    uint32_t* scount = reinterpret_cast<uint32_t*>(sarea + k);         // [2][kMaxFrameLists] list lengths to stream / full lengths
    WaveShared* wsh = reinterpret_cast<WaveShared*>(scount + 2 * kMaxFrameLists);
    lists += (size_t)f * n_lists * kListCap; counts += (size_t)f * n_lists; counts_hi += (size_t)f * n_lists; hist += (size_t)f * kHistBins;

    VD_STAMP(P, 0);
    if (P.dbg && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
        unsigned int sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        P.dbg[8192 + f * 4 + 0] = (long long)t; P.dbg[8192 + f * 4 + 2] = (long long)sm;
    }
    if (tid < 64) coarse[(size_t)f * 64 + tid] = 0u;
    select_scratch_init(scr);
    // ---- 1. suffix scan of the histogram; thread t owns bins [4095-16t-15, 4095-16t], highest first.
    //         The bins are zeroed behind the read: the next call's head kernel accumulates from zero.
    constexpr int BPT = kHistBins / kNmsThreads;
    uint32_t h[BPT], local = 0u;
    {
        uint4* h4 = reinterpret_cast<uint4*>(hist + kHistBins - (tid + 1) * BPT);   // ascending bins of this thread's range
#pragma unroll
        for (int i = 0; i < BPT / 4; ++i) {
            const uint4 v = h4[i];
            h[BPT - 1 - 4 * i] = v.x; h[BPT - 2 - 4 * i] = v.y; h[BPT - 3 - 4 * i] = v.z; h[BPT - 4 - 4 * i] = v.w;
            h4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < BPT; ++i) local += h[i];
    }
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) sx[warp] = incl;
    __syncthreads();
    uint32_t woff = 0u, total = 0u;
    for (int w = 0; w < nwarps; ++w) { uint32_t v = sx[w]; if (w < warp) woff += v; total += v; }
    if (tid == 0) { sx[32] = 0u; sx[33] = total; }                     // defaults: fewer than k candidates -> take all
    __syncthreads();
    const uint32_t excl = woff + incl - local;
    if (excl < (uint32_t)k && excl + local >= (uint32_t)k) {
        uint32_t run = excl;
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            run += h[i];
            if (run >= (uint32_t)k) { sx[32] = (uint32_t)(kHistBins - 1 - (tid * BPT + i)); sx[33] = run; break; }
        }
    }
    __syncthreads();
    VD_STAMP(P, 1);
    uint32_t bstar = sx[32];
    uint32_t cnt = sx[33];
    uint64_t piv = (uint64_t)hist_edge(bstar) << 32;

    auto stream_count = [&](uint64_t pv, int& it) -> uint32_t {
        uint32_t c = 0u;
        for (int l = warp; l < n_lists; l += nwarps) {
            uint32_t n = counts[l]; n = n > (uint32_t)kListCap ? (uint32_t)kListCap : n;
            for (uint32_t j = lane; j < n; j += 32) { uint64_t v = lists[(size_t)l * kListCap + j]; c += (v >= pv && v != 0ull) ? 1u : 0u; }
        }
        return block_sum(c, scr, it);
    };
    // lists are split at hint_hi: every key >= hint_hi sits in the first counts_hi entries.  A pivot at or above the
    // hint needs only those fronts; the hint for the next call trails this frame's pivot by 16 bins (~2 % in score).
    const uint32_t hint_used = hint_hi[f];
    const bool fronts_only = (uint32_t)(piv >> 32) >= hint_used && cnt <= (uint32_t)kHistCap;
    for (int l = tid; l < n_lists; l += blockDim.x) {
        uint32_t n = counts[l]; n = n > (uint32_t)kListCap ? (uint32_t)kListCap : n;
        uint32_t nh = counts_hi[l]; nh = nh > n ? n : nh;
        scount[l] = fronts_only ? nh : n;
        if (fronts_only) scount[kMaxFrameLists + l] = n;
    }
    if (tid == 0) hint_hi[f] = hist_edge(bstar > 16u ? bstar - 16u : 0u);
    int sit = 0;
    uint32_t m = 0u;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (cnt > (uint32_t)kHistCap || attempt == 1) {
            // rare: one bin holds a huge tie cluster, or (attempt 1) the histogram did not describe the lists
            // (foreign workspace content): exact streaming bisection on the 64-bit keys
            uint64_t lo = attempt == 1 ? 1ull : piv;
            uint64_t hi = (attempt == 0 && bstar + 1u < (uint32_t)kHistBins) ? ((uint64_t)hist_edge(bstar + 1u) << 32) : ~0ull;
            uint32_t c = stream_count(lo, sit);
            if (c > (uint32_t)kHistCap) {
                for (int g = 0; g < 80 && hi - lo > 1ull; ++g) {
                    const uint64_t mid = lo + ((hi - lo) >> 1);
                    c = stream_count(mid, sit);
                    if (c > (uint32_t)kHistCap) lo = mid; else if (c < (uint32_t)k) hi = mid; else { lo = mid; break; }
                }
            }
            piv = lo;
        }
        // ---- 2. stream the tile lists once, keep keys >= pivot; 8 independent 8-byte loads per lane in flight
        __syncthreads();
        if (tid == 0) scr->out_count = 0u;
        __syncthreads();
        for (int l = warp; l < n_lists; l += nwarps) {
            const uint32_t nl = scount[l];
            const uint64_t* src_l = lists + (size_t)l * kListCap;
            for (uint32_t b0 = 0; b0 < nl; b0 += 256u) {
                uint64_t v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const uint32_t j = b0 + (uint32_t)(u * 32 + lane); v[u] = (j < nl) ? src_l[j] : 0ull; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool keep = v[u] >= piv && v[u] != 0ull;
                    const unsigned mm = __ballot_sync(0xffffffffu, keep);
                    if (mm) {
                        uint32_t base = 0u;
                        if (lane == 0) base = atomicAdd(&scr->out_count, (uint32_t)__popc(mm));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const uint32_t pos = base + __popc(mm & ((1u << lane) - 1u));
                        if (keep && pos < (uint32_t)kHistCap) skeys[pos] = v[u];
                    }
                }
            }
        }
        __syncthreads();
        m = scr->out_count;
        // a clean histogram guarantees m >= k whenever a positive pivot was chosen; otherwise redo exactly
        if (m >= (uint32_t)k || piv <= 1ull || attempt == 1) break;
        if (fronts_only) { for (int l = tid; l < n_lists; l += blockDim.x) scount[l] = scount[kMaxFrameLists + l]; }   // the exact retry reads whole lists
    }
    VD_STAMP(P, 2);
    m = m > (uint32_t)kHistCap ? (uint32_t)kHistCap : m;
    const int SN = m <= 512u ? 512 : 1024;
    for (int i = (int)m + tid; i < SN; i += blockDim.x) skeys[i] = 0ull;
    __syncthreads();
    block_sort_u64_desc(skeys, SN);
    VD_STAMP(P, 3);
    const int n = (int)(m < (uint32_t)k ? m : (uint32_t)k);
    nms_tail_wave(n, f, P, src, sink, skeys, sbox, scls, sarea, skbox, skarea, skcls, wsh);
    if (spec_tau && tid == 0) spec_tau[f] = hist_edge(bstar > 16u ? bstar - 16u : 0u) & 0x7fffffffu;   // the next call's threshold for this frame slot
    if (P.dbg && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); P.dbg[8192 + f * 4 + 1] = (long long)t; }
    }   // active
    if (spec_state == nullptr) {
        if (blockIdx.x == 0 && tid == 0) { ctr[0] = 0u; ctr[1] = 0u; ctr[2] = ws_magic; }   // workspace is in its between-calls state (for this layout)
        return;
    }
    // last CTA out closes the call: counters re-armed, workspace marked as being in its between-calls state
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&spec_state[3], 1u) == gridDim.x - 1u) {
            spec_state[4] = spec_state[2]; spec_state[5] += spec_state[2]; spec_state[6] += 1u;     // last call / running totals (frames redone, calls)
            spec_state[2] = 0u; spec_state[3] = 0u;
            ctr[0] = 0u; ctr[1] = 0u; ctr[2] = ws_magic;
        }
    }
}

// Speculative path, per frame: the head kernel appended every candidate that can reach the call's threshold tau to the
// frame's list.  The list is the frame's exact top-k source iff it did not overflow and holds >= k keys at or above tau
// (or tau is the valid floor: then it simply holds every valid candidate).  Such frames are finished here (sort, wavefront
// NMS, outputs); the others are queued for the exact path.  256 threads, 56 registers: shares SMs with a head kernel.
__global__ void __maxnreg__(56)
nms_spec_kernel(const uint64_t* __restrict__ spec_lists, uint32_t* __restrict__ spec_cnt, uint32_t* __restrict__ spec_state, uint32_t* __restrict__ spec_tau,
                uint32_t* __restrict__ failed, const unsigned int* __restrict__ ctr, unsigned int ws_magic, float valid_thresh,
                NmsParams P, FusedSource src, FusedSink sink) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int k = P.k;
    SelectScratch* scr = reinterpret_cast<SelectScratch*>(smem_raw);
    uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + 256);
    float4* sbox = reinterpret_cast<float4*>(skeys + kSpecCap);
    const int KMAX = (P.max_out < k ? P.max_out : k) + 32 + 4 * (kNmsThreads / 32);
    float4* skbox = sbox + k;
    float* skarea = reinterpret_cast<float*>(skbox + KMAX);
    int* skcls = reinterpret_cast<int*>(skarea + KMAX);
    int* scls = skcls + KMAX;
    float* sarea = reinterpret_cast<float*>(scls + k);
    WaveShared* wsh = reinterpret_cast<WaveShared*>(sarea + k);
    VD_STAMP(P, 0);
    if (P.dbg && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
        unsigned int sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        P.dbg[8192 + f * 4 + 0] = (long long)t; P.dbg[8192 + f * 4 + 2] = (long long)sm;
    }
    select_scratch_init(scr);
    const bool ws_ok = ctr[2] == ws_magic;
    const uint32_t cnt = spec_cnt[f];
    const uint32_t floor_b = valid_thresh > 0.0f ? __float_as_uint(valid_thresh) : 0u;
    uint32_t tb = ws_ok ? spec_tau[f] : 0u;
    tb = tb > floor_b ? tb : floor_b;
    if (tb > 0x3f800001u) tb = 0x3f800001u;
    __syncthreads();
    if (tid == 0) spec_cnt[f] = 0u;                              // next call appends from zero
    if (!ws_ok) {
        // foreign workspace: nothing in it can be trusted -> every frame takes the exact path, which also re-initialises it
        if (tid == 0) {
            failed[f] = (uint32_t)f;
            if (f == 0) { spec_state[2] = gridDim.x; spec_state[3] = 0u; }
        }
        return;
    }
    bool ok = cnt <= (uint32_t)kSpecCap;
    uint32_t n_ge = 0u;
    if (ok) {
        const uint32_t tk = tb | 0x80000000u;
        uint32_t c = 0u;
        for (uint32_t j = tid; j < cnt; j += blockDim.x) {
            const uint64_t v = spec_lists[(size_t)f * kSpecCap + j];
            skeys[j] = v;
            c += ((uint32_t)(v >> 32) >= tk) ? 1u : 0u;
        }
        int it = 0;
        n_ge = block_sum(c, scr, it);
        ok = n_ge >= (uint32_t)k || tb == floor_b;
    }
    if (!ok) {
        if (tid == 0) { const uint32_t slot_ = atomicAdd(&spec_state[2], 1u); VD_DEV_CHECK(slot_ < gridDim.x); failed[slot_] = (uint32_t)f; }
        return;
    }
    VD_STAMP(P, 2);
    static_assert(kSpecCap == 1024 || kSpecCap == 2048, "kSpecCap");
    const int SN = cnt <= 512u ? 512 : (cnt <= 1024u ? 1024 : 2048);
    for (int i = (int)cnt + tid; i < SN; i += blockDim.x) skeys[i] = 0ull;
    __syncthreads();
    block_sort_u64_desc(skeys, SN);
    VD_STAMP(P, 3);
    const int n = (int)(cnt < (uint32_t)k ? cnt : (uint32_t)k);
    // the next call's threshold for this frame slot: the score at rank ~2 k of this frame, a little lower
    if (tid == 0) {
        uint32_t d = floor_b;
        if (cnt >= (uint32_t)k) {
            // aim at ~2k candidates (1.5k when the list holds <= 1024): room for the next frame in this slot to differ both ways
#ifndef VD_SPEC_TARGET_PCT
#define VD_SPEC_TARGET_PCT (kSpecCap >= 2048 ? 200 : 150)
#endif
            const uint32_t r = min(cnt - 1u, (uint32_t)((VD_SPEC_TARGET_PCT * k) / 100));
            const uint32_t bits = (uint32_t)(skeys[r] >> 32) & 0x7fffffffu;
            d = bits > floor_b + 65536u ? bits - 65536u : floor_b;          // ~0.8 % lower
        }
        spec_tau[f] = d;
    }
    nms_tail_wave(n, f, P, src, sink, skeys, sbox, scls, sarea, skbox, skarea, skcls, wsh);
    if (P.dbg && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); P.dbg[8192 + f * 4 + 1] = (long long)t; }
}
static size_t nms_spec_smem(int k, int max_out) {
    const size_t kmax = (size_t)(max_out < k ? max_out : k) + 32 + 4 * (kNmsThreads / 32);
    return 256 + (size_t)kSpecCap * 8 + (size_t)k * 16 + kmax * 24 + (size_t)k * 8 + sizeof(WaveShared) + 64;
}
static size_t nms_hist_smem(int k, int max_out) {
    const size_t kmax = (size_t)(max_out < k ? max_out : k) + 32 + 4 * (kNmsThreads / 32);
    return 256 + (size_t)kHistCap * 8 + (size_t)k * 16 + kmax * 24 + (size_t)k * 8 + 2 * kMaxFrameLists * 4 + sizeof(WaveShared) + 64;
}

// no-NMS tail (yolo3.py:525 false): rows are the plain concat; only reachable through vd_head_detections.

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
struct HeadPlan {
    HeadKernelParams kp;
    int n_pad, C;                  // C: class count of the compiled kernel shape that runs (== num_class when that is compiled)
    int C_total, n_pass;           // the head's num_class; class windows of C classes each
    bool repack;                   // weights / biases are re-laid out per window into the workspace (off_wrepack)
    size_t off_wrepack, wrepack_bytes, wrepack_w_bytes[VD_MAX_SCALES], wrepack_stride[VD_MAX_SCALES];
    int merge_levels;
    size_t off_hints, off_ctr, off_hist, off_boxes, off_lists0, off_counts0, off_counts_hi, off_hint_hi, off_coarse, off_spec_lists, off_spec_cnt, off_spec_state, off_spec_tau, off_failed, off_winlist, off_listsA, off_listsB, off_countsA, off_countsB, total;
};

}  // namespace vd
#include "tfused.cuh"
#include "hpair.cuh"
namespace vd {

static int head_npad(int C) { int n = 3 * (5 + C); return (n + 15) / 16 * 16; }

// Plane products of the fp32-parity modes.  v = p0 + p1 (+ p2) with |p1| <= 2^-9 |v|, |p2| <= 2^-18 |v|.
//   3 planes (VD_PREC_FP32_SPLIT):  p0 w2 + p2 w0 + p1 w1 + p0 w1 + p1 w0 + p0 w0   (dropped terms <= 2^-27 relative)
//   2 planes (VD_PREC_BF16X2):      p0 w1 + p1 w0 + p0 w0                           (dropped / residual terms ~ 2^-18 relative)
// The tensor core's fp32 accumulator truncates on every accumulate (measured: the error of a K-term bf16 dot product grows
// LINEARLY, ~2.3e-8 of max|y| per MMA, tests/test_gpu_head.py prints it).  So in these modes (a) the low-order products are
// summed in their own accumulator (their truncation errors scale with their own 2^-9 / 2^-18 magnitude), (b) the p0 w0 chain
// is cut into up to 3 K-ranges with separate accumulators, and (c) the epilogue adds the partial accumulators in fp32
// round-to-nearest before decoding (fold, see head_kernel).  The last product listed here is p0 w0.
static int split_products(int precision, HeadKernelParams* k) {
    static const signed char a3[6] = {0, 2, 1, 0, 1, 0}, w3[6] = {2, 0, 1, 1, 0, 0};
    static const signed char a2[3] = {0, 1, 0}, w2[3] = {1, 0, 0};
    const bool three = precision == VD_PREC_FP32_SPLIT;
    const int n = three ? 6 : 3;
    k->split = three ? 3 : 2;
    k->K_frames = n;
    for (int i = 0; i < 8; ++i) { k->a_pl[i] = i < n ? (three ? a3[i] : a2[i]) : 0; k->w_pl[i] = i < n ? (three ? w3[i] : w2[i]) : 0; }
    return n;
}

static int make_plan(const VdHeadParams* hp, HeadPlan* pl) {
    VD_CHECK_ARG(hp, "head: null params");
    VD_CHECK_ARG(hp->num_scales >= 1 && hp->num_scales <= VD_MAX_SCALES, "head: num_scales %d", hp->num_scales);
    VD_CHECK_ARG(hp->num_class >= 1, "head: num_class %d", hp->num_class);
    VD_CHECK_ARG(hp->frames >= 0 && hp->frames <= 65535, "head: frames %d out of range", hp->frames);
    // Any num_class runs on the compiled kernel shapes (C in 1-5, 20, 30, 80): a class count in between is padded up to the next
    // shape, a larger one is cut into windows of 80 classes (one head-kernel launch per window, all appending to the same
    // per-frame candidate lists).  The prediction weights / biases of a window are re-laid out into the workspace first
    // (repack_pred_weights_kernel): box + objectness rows of each anchor, then the window's class rows, padding rows zero.
    const int C_total = hp->num_class;
    int C = C_total, n_pass = 1;
    bool repack = false;
    if (!(C_total <= 5 || C_total == 20 || C_total == 30 || C_total == 80)) {
        repack = true;
        C = C_total < 20 ? 20 : (C_total < 30 ? 30 : 80);
        n_pass = ceil_div(C_total, C);
    }
    const int npad = head_npad(C);
    memset(pl, 0, sizeof(*pl));
    HeadKernelParams& k = pl->kp;
    pl->n_pad = npad; pl->C = C; pl->C_total = C_total; pl->n_pass = n_pass; pl->repack = repack;
    k.g.num_scales = hp->num_scales; k.g.num_class = C_total; k.g.A = 3;
    k.frames = hp->frames; k.n_pad = npad; k.n_valid = 3 * (5 + C);
    k.c_off = 0; k.c_valid = C_total < C ? C_total : C; k.n_pass = n_pass; k.pass = 0;
    const int join = hp->join;
    VD_CHECK_ARG(join == VD_JOIN_NONE || join == VD_JOIN_CAT, "head: join %d must be pre-reduced (use vd_temporal_pool for max/mean)", join);
    VD_CHECK_ARG(hp->precision == VD_PREC_BF16 || hp->precision == VD_PREC_FP32_SPLIT || hp->precision == VD_PREC_BF16X2, "head: precision %d", hp->precision);
    k.K_frames = (join == VD_JOIN_CAT) ? hp->K_frames : 1;
    VD_CHECK_ARG(k.K_frames >= 1, "head: K_frames %d", k.K_frames);
    if (hp->precision != VD_PREC_BF16) {
        if (join != VD_JOIN_NONE) return set_error(VD_ERR_UNSUPPORTED, "head: the fp32-parity modes are not combinable with a late cat join");
        for (int s = 0; s < hp->num_scales; ++s)
            if (hp->scale[s].tconv_weight_bf16) return set_error(VD_ERR_UNSUPPORTED, "head: the fp32-parity modes are not combinable with the fused temporal tip cell (run vd_temporal_conv_ex first)");
        split_products(hp->precision, &k);
    }
    int rows = 0, anc = 0, tif = 0;
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        VD_CHECK_ARG(sc.H > 0 && sc.W > 0 && sc.H <= 128 && sc.W <= 128, "head: scale %d feature map %dx%d (alloc_size is 128x128, yolo3.py:44)", s, sc.H, sc.W);
        VD_CHECK_ARG(sc.Cin > 0 && sc.Cin % BLOCK_K == 0, "head: scale %d Cin %d must be a multiple of %d", s, sc.Cin, BLOCK_K);
        VD_CHECK_ARG((sc.tip_nhwc_bf16 || hp->frames == 0) && sc.weight_bf16, "head: scale %d null tip/weight", s);     // an empty batch has no tip storage
        VD_CHECK_ARG(((uintptr_t)sc.tip_nhwc_bf16 & 15) == 0 && ((uintptr_t)sc.weight_bf16 & 15) == 0, "head: scale %d tensors must be 16-byte aligned", s);
        k.g.H[s] = sc.H; k.g.W[s] = sc.W; k.g.HW[s] = sc.H * sc.W;
        k.g.stride[s] = sc.stride;
        for (int i = 0; i < 6; ++i) k.g.anchors[s][i] = sc.anchors[i];
        k.g.row_base[s] = rows; k.g.anc_base[s] = anc;
        rows += C_total * k.g.HW[s] * 3; anc += k.g.HW[s] * 3;
        k.cin[s] = sc.Cin; k.bias[s] = sc.bias;
        k.pb[s] = ceil_div(k.g.HW[s], BLOCK_M);
        k.tif_base[s] = tif; tif += k.pb[s];
    }
    // processing order: largest K (most bytes per tile) first, smallest tiles last -- the dynamic scheduler's tail is one small tile
    int ord[VD_MAX_SCALES] = {0, 1, 2};
    for (int a = 0; a < hp->num_scales; ++a)
        for (int b = a + 1; b < hp->num_scales; ++b)
            if (k.cin[ord[b]] > k.cin[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    int tiles = 0;
    for (int j = 0; j < hp->num_scales; ++j) { k.order[j] = ord[j]; k.tile_start[j] = tiles; tiles += k.pb[ord[j]] * hp->frames; }
    for (int j = hp->num_scales; j < VD_MAX_SCALES; ++j) k.order[j] = 0;
    for (int s = hp->num_scales; s <= VD_MAX_SCALES; ++s) { k.g.row_base[s] = rows; k.g.anc_base[s] = anc; k.tile_start[s] = tiles; }
    k.tiles_per_frame = tif; k.total_tiles = tiles;
    if (tif * n_pass > kMaxFrameLists) return set_error(VD_ERR_UNSUPPORTED, "head: %d pixel blocks x %d class windows per frame > %d", tif, n_pass, kMaxFrameLists);
    // workspace
    size_t off = 0;
    const size_t F = (size_t)(hp->frames > 0 ? hp->frames : 1);
    pl->off_hints = off; off += 256;
    pl->off_ctr = off; off += 256;                      // dynamic tile counter; zeroed together with the histogram that follows it
    pl->off_hist = off; off += align_up(F * kHistBins * 4, 256);
    pl->off_boxes = off; off += align_up(F * anc * 16, 256);
    pl->off_lists0 = off; off += align_up(F * tif * n_pass * kListCap * 8, 256);
    pl->off_counts0 = off; off += align_up(F * tif * n_pass * 4, 256);
    pl->off_counts_hi = off; off += align_up(F * tif * n_pass * 4, 256);
    pl->off_hint_hi = off; off += align_up(F * 4, 256);
    pl->off_coarse = off; off += align_up(F * 64 * 4, 256);
    pl->off_spec_lists = off; off += align_up(F * kSpecCap * 8, 256);
    pl->off_spec_cnt = off; off += align_up(F * 4, 256);
    pl->off_spec_state = off; off += 256;
    pl->off_spec_tau = off; off += align_up(F * 4, 256);
    pl->off_failed = off; off += align_up(F * 4, 256);
    pl->off_winlist = off; off += align_up((F + 64) * 4, 256);       // fused temporal head: windows of the failed frames ([0] = count, [64..] = first frame of each window)
    int n1 = ceil_div(tif, kMaxLists);
    pl->off_listsA = off; off += align_up(F * n1 * kListCap * 8, 256);
    pl->off_listsB = off; off += align_up(F * n1 * kListCap * 8, 256);
    pl->off_countsA = off; off += align_up(F * n1 * 4, 256);
    pl->off_countsB = off; off += align_up(F * n1 * 4, 256);
    pl->off_wrepack = off;
    if (repack) {
        const int planes = hp->precision == VD_PREC_FP32_SPLIT ? 3 : (hp->precision == VD_PREC_BF16X2 ? 2 : 1);
        for (int s = 0; s < hp->num_scales; ++s) {
            const size_t row_elems = (size_t)hp->scale[s].Cin * (planes > 1 ? planes : k.K_frames);
            pl->wrepack_w_bytes[s] = align_up((size_t)npad * row_elems * 2, 256);
            pl->wrepack_stride[s] = pl->wrepack_w_bytes[s] + align_up((size_t)npad * 4, 256);     // weights then biases of one window
            off += pl->wrepack_stride[s] * n_pass;
        }
    }
    pl->wrepack_bytes = off - pl->off_wrepack;
    pl->total = off;
    return VD_OK;
}

// Weights / biases of class window `pass` in the layout of the compiled shape (C_T classes): row a*(5+C_T)+p <- source row
// a*(5+C_total)+p for the box / objectness part (p < 5), a*(5+C_total)+5+c_off+(p-5) for the window's classes; padding rows zero.
__global__ void __launch_bounds__(256)
repack_pred_weights_kernel(const uint4* __restrict__ w, const float* __restrict__ bias, uint4* __restrict__ w_out, float* __restrict__ b_out,
                           int C_total, int C_T, int c_off, int npad, int row_vec /* 16-byte vectors per row */) {
    const int n = blockIdx.x;                  // output row
    const int P_T = 5 + C_T, P_S = 5 + C_total;
    const int a = n / P_T, pp = n - a * P_T;
    int src = -1;
    if (a < 3) {
        if (pp < 5) src = a * P_S + pp;
        else { const int c = c_off + pp - 5; if (c < C_total) src = a * P_S + 5 + c; }
    }
    for (int j = threadIdx.x; j < row_vec; j += blockDim.x)
        w_out[(size_t)n * row_vec + j] = src >= 0 ? w[(size_t)src * row_vec + j] : make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) b_out[n] = (src >= 0 && bias) ? bias[src] : 0.0f;
    (void)npad;
}

// Runs the repack for every (scale, window) and returns, through wptr / bptr, where window `pass` of scale s lives.
static int repack_windows(const VdHeadParams* hp, const HeadPlan& pl, unsigned char* ws, cudaStream_t stream) {
    if (!pl.repack) return VD_OK;
    size_t off = pl.off_wrepack;
    const int planes = hp->precision == VD_PREC_FP32_SPLIT ? 3 : (hp->precision == VD_PREC_BF16X2 ? 2 : 1);
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        const size_t row_elems = (size_t)sc.Cin * (planes > 1 ? planes : pl.kp.K_frames);
        for (int ps = 0; ps < pl.n_pass; ++ps) {
            unsigned char* base = ws + off + (size_t)ps * pl.wrepack_stride[s];
            repack_pred_weights_kernel<<<pl.n_pad, 256, 0, stream>>>((const uint4*)sc.weight_bf16, sc.bias, (uint4*)base, (float*)(base + pl.wrepack_w_bytes[s]),
                                                                      pl.C_total, pl.C, ps * pl.C, pl.n_pad, (int)(row_elems / 8));
            VD_LAUNCH_CHECK();
        }
        off += pl.wrepack_stride[s] * pl.n_pass;
    }
    return VD_OK;
}
// Kernel parameters / weight pointers of class window `pass`.
static void window_params(const VdHeadParams* hp, const HeadPlan& pl, unsigned char* ws, int pass, HeadKernelParams* kp, const void* wptr[VD_MAX_SCALES]) {
    kp->pass = pass; kp->n_pass = pl.n_pass; kp->c_off = pass * pl.C;
    const int left = pl.C_total - kp->c_off;
    kp->c_valid = left < pl.C ? left : pl.C;
    size_t off = pl.off_wrepack;
    for (int s = 0; s < hp->num_scales; ++s) {
        if (pl.repack) {
            unsigned char* base = ws + off + (size_t)pass * pl.wrepack_stride[s];
            wptr[s] = base; kp->bias[s] = (const float*)(base + pl.wrepack_w_bytes[s]);
            off += pl.wrepack_stride[s] * pl.n_pass;
        } else { wptr[s] = hp->scale[s].weight_bf16; kp->bias[s] = hp->scale[s].bias; }
    }
}

static int make_maps(const VdHeadParams* hp, const HeadPlan& pl, HeadMaps* maps, const void* const* wptr = nullptr) {
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        const uint64_t HW = (uint64_t)sc.H * sc.W, Cin = sc.Cin, K = pl.kp.split ? pl.kp.split : pl.kp.K_frames;
        const uint64_t F = (uint64_t)(hp->frames > 0 ? hp->frames : 1);
        const void* aptr = (sc.tconv_weight_bf16 && sc.tconv_out_nhwc_bf16) ? sc.tconv_out_nhwc_bf16 : sc.tip_nhwc_bf16;
        uint64_t dimsA[4] = {Cin, HW, K, F};
        uint64_t strA[3] = {Cin * 2, HW * Cin * 2, K * HW * Cin * 2};
        if (pl.kp.split) { strA[1] = F * HW * Cin * 2; strA[2] = HW * Cin * 2; }     // plane-major carrier (planes, frames, H, W, Cin): every plane is an ordinary NHWC tensor
        uint32_t boxA[4] = {BLOCK_K, BLOCK_M, 1, 1};
        int rc = encode_tmap_bf16(&maps->a[s], aptr, 4, dimsA, strA, boxA);
        if (rc) return rc;
        uint64_t dimsW[2] = {Cin * K, (uint64_t)3 * (5 + pl.C)};
        uint64_t strW[1] = {Cin * K * 2};
        uint32_t boxW[2] = {BLOCK_K, (uint32_t)pl.n_pad};
        rc = encode_tmap_bf16(&maps->w[s], wptr ? wptr[s] : sc.weight_bf16, 2, dimsW, strW, boxW);
        if (rc) return rc;
    }
    return VD_OK;
}

template <int EPI, int C, int NPAD>
static int launch_head_t(const HeadMaps& maps, const HeadKernelParams& kp, cudaStream_t stream) {
    using Cfg = HeadCfg<EPI, C, NPAD>;
    auto kern = head_kernel<EPI, C, NPAD>;
    // once per (device, instantiation) (also keeps graph capture clean); same carve-out as the NMS kernel: CTAs of both can share an SM
    { int rc_ = configure_kernel((const void*)kern, Cfg::SMEM_BYTES, !getenv("VD_DEBUG_NO_CARVEOUT")); if (rc_) return rc_; }
    int grid = sm_count(); if (grid > kp.total_tiles) grid = kp.total_tiles;
    // the exact fallback of the speculative path is idle in the steady state: a handful of CTAs slip in between two
    // head kernels instead of claiming every SM's shared memory (when frames did fail, they work through them slowly)
    if (kp.frame_list && grid > kFallbackCtas) grid = kFallbackCtas;
    if (grid < 1) return VD_OK;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = getenv("VD_PDL") ? 1 : 0;   // measured slower in the pipeline (43.2 vs 38.6 us/step): off unless asked for
    cfg.attrs = attr; cfg.numAttrs = 1;
    VD_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, kp));
    VD_LAUNCH_CHECK();
    return VD_OK;
}

template <int EPI>
static int launch_head(const HeadMaps& maps, const HeadKernelParams& kp, int C, cudaStream_t stream) {
    switch (C) {
        case 20: return launch_head_t<EPI, 20, 80>(maps, kp, stream);     // VOC
        case 30: return launch_head_t<EPI, 30, 112>(maps, kp, stream);    // ImageNet-VID
        case 80: return launch_head_t<EPI, 80, 256>(maps, kp, stream);    // COCO
        case 1:  return launch_head_t<EPI, 1, 32>(maps, kp, stream);
        case 2:  return launch_head_t<EPI, 2, 32>(maps, kp, stream);
        case 3:  return launch_head_t<EPI, 3, 32>(maps, kp, stream);
        case 4:  return launch_head_t<EPI, 4, 32>(maps, kp, stream);
        case 5:  return launch_head_t<EPI, 5, 32>(maps, kp, stream);
        default: break;
    }
    return set_error(VD_ERR_UNSUPPORTED, "head: internal error, no kernel shape for %d classes", C);
}

// EPI_PRED only depends on the padded width
static int launch_pred(const HeadMaps& maps, const HeadKernelParams& kp, cudaStream_t stream) {
    switch (kp.n_pad) {
        case 16:  return launch_head_t<EPI_PRED, 1, 16>(maps, kp, stream);
        case 32:  return launch_head_t<EPI_PRED, 1, 32>(maps, kp, stream);
        case 48:  return launch_head_t<EPI_PRED, 1, 48>(maps, kp, stream);
        case 64:  return launch_head_t<EPI_PRED, 1, 64>(maps, kp, stream);
        case 80:  return launch_head_t<EPI_PRED, 1, 80>(maps, kp, stream);
        case 96:  return launch_head_t<EPI_PRED, 1, 96>(maps, kp, stream);
        case 112: return launch_head_t<EPI_PRED, 1, 112>(maps, kp, stream);
        case 128: return launch_head_t<EPI_PRED, 1, 128>(maps, kp, stream);
        case 144: return launch_head_t<EPI_PRED, 1, 144>(maps, kp, stream);
        case 160: return launch_head_t<EPI_PRED, 1, 160>(maps, kp, stream);
        case 176: return launch_head_t<EPI_PRED, 1, 176>(maps, kp, stream);
        case 192: return launch_head_t<EPI_PRED, 1, 192>(maps, kp, stream);
        case 208: return launch_head_t<EPI_PRED, 1, 208>(maps, kp, stream);
        case 224: return launch_head_t<EPI_PRED, 1, 224>(maps, kp, stream);
        case 240: return launch_head_t<EPI_PRED, 1, 240>(maps, kp, stream);
        case 256: return launch_head_t<EPI_PRED, 1, 256>(maps, kp, stream);
        default: break;
    }
    return set_error(VD_ERR_INVALID_ARG, "pred_conv: bad padded width %d", kp.n_pad);
}

// Wide heads (256 prediction columns) run the speculative head kernel on CTA pairs (hpair.cuh) where it applies.
static bool hpair_applicable(const VdHeadParams* hp, const HeadPlan& pl, bool spec) {
    static const bool on = []() { const char* e = getenv("VD_HEAD_PAIR"); return e ? atoi(e) != 0 : true; }();
    if (!on || !spec || (hp->flags & VD_HEAD_NO_PAIR_KERNEL) || pl.n_pass != 1 || hp->precision != VD_PREC_BF16 || hp->join != VD_JOIN_NONE || pl.kp.K_frames != 1) return false;
    if (!hpair_supported(pl.C) || hp->frames <= 0) return false;
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        if (sc.tconv_weight_bf16) return false;
        const long long tiles = ceil_div((long long)hp->frames * sc.H * sc.W, 128);
        if (tiles < 2 || tiles > 2 * 65000) return false;
    }
    return true;
}

static int run_hpair(const VdHeadParams* hp, const HeadPlan& pl, const HeadKernelParams& kp, const void* const* wptr, cudaStream_t stream) {
    PairParams pp;
    PairMaps maps;
    memset(&pp, 0, sizeof(pp));
    memset(&maps, 0, sizeof(maps));
    pp.num_scales = hp->num_scales; pp.frames = hp->frames;
    pp.g = kp.g; pp.c_valid = kp.c_valid; pp.valid_thresh = hp->valid_thresh;
    pp.boxes = kp.boxes; pp.spec_lists = kp.spec_lists; pp.spec_cnt = kp.spec_cnt; pp.spec_tau = kp.spec_tau;
    pp.tile_counter = kp.tile_counter; pp.ws_magic = kp.ws_magic;
    { const char* e = getenv("VD_HEAD_PAIR_DBG"); pp.dbg = e ? atoi(e) : 0; }
    int clusters = sm_count() / 2;
    if (clusters > F_MAX_CLUSTERS) clusters = F_MAX_CLUSTERS;
    pp.pairs = clusters;
    double load[F_MAX_CLUSTERS];
    int cnt[F_MAX_CLUSTERS];
    for (int c = 0; c < clusters; ++c) load[c] = 0.0;
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        PairScale& q = pp.sc[s];
        q.HW = sc.H * sc.W; q.Cin = sc.Cin; q.rows = hp->frames * sc.H * sc.W; q.m_tiles = ceil_div(q.rows, 128); q.bias = kp.bias[s];
        const uint64_t Cin = (uint64_t)sc.Cin;
        uint64_t dimsA[2] = {Cin, (uint64_t)q.rows};
        uint64_t strA[1] = {Cin * 2};
        uint32_t boxA[2] = {64, 128};
        int rc = encode_tmap_bf16(&maps.a[s], sc.tip_nhwc_bf16, 2, dimsA, strA, boxA);
        if (rc) return rc;
        uint64_t dimsW[2] = {Cin, (uint64_t)3 * (5 + pl.C)};
        uint32_t boxW[2] = {64, (uint32_t)(pl.n_pad / 2)};
        rc = encode_tmap_bf16(&maps.w[s], wptr[s], 2, dimsW, strA, boxW);
        if (rc) return rc;
        // scales in the order given (deep -> shallow = decreasing K): an item costs its k-blocks + the decode of two tiles
        const int items = (q.m_tiles + 1) / 2;
        lpt_fill(load, clusters, items, 0.55 * (sc.Cin / 64) + 3.0, cnt);
        int acc = 0;
        for (int c = 0; c < clusters; ++c) { pp.beg[s][c] = (unsigned short)acc; acc += cnt[c]; }
        for (int c = clusters; c <= F_MAX_CLUSTERS; ++c) pp.beg[s][c] = (unsigned short)acc;
    }
    return launch_hpair(maps, pp, pl.C, clusters, stream);
}

// Exact fallback of the fused temporal head: the failed FRAMES (queued by nms_spec_kernel) -> the distinct WINDOWS they belong to, so that
// the conditional tip-cell launches recompute every such window once (on a cold workspace all T frames of every window are listed).
// out[0] = number of windows, out[64 + i] = first frame of window i.  One CTA; a bitmap of the windows in shared memory.
__global__ void __launch_bounds__(1024)
failed_windows_kernel(const uint32_t* __restrict__ failed, const uint32_t* __restrict__ n_failed, int T, int n_windows, uint32_t* __restrict__ out) {
    __shared__ uint32_t bits[2048];                    // up to 65 536 windows
    __shared__ uint32_t n_out;
    const uint32_t n = *n_failed;
    if (n == 0u) { if (threadIdx.x == 0) out[0] = 0u; return; }
    const int words = (n_windows + 31) >> 5;
    for (int i = threadIdx.x; i < words; i += blockDim.x) bits[i] = 0u;
    if (threadIdx.x == 0) n_out = 0u;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { const uint32_t w = failed[i] / (uint32_t)T; atomicOr(&bits[w >> 5], 1u << (w & 31u)); }
    __syncthreads();
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
        uint32_t m = bits[i];
        while (m) { const int b = __ffs(m) - 1; m &= m - 1u; out[64u + atomicAdd(&n_out, 1u)] = (uint32_t)((i * 32 + b) * T); }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[0] = n_out;
}

// The fused tip-cell + head kernel (tfused.cuh) replaces the temporal_conv + head_kernel<EPI_SPEC> launches where it applies.
static bool tfused_applicable(const VdHeadParams* hp, const HeadPlan& pl, bool spec) {
    static const bool on = []() { const char* e = getenv("VD_TFUSED"); return e ? atoi(e) != 0 : true; }();
    if (!on || !spec || (hp->flags & VD_HEAD_NO_FUSED_TIP)) return false;
    if (pl.n_pass != 1 || hp->precision != VD_PREC_BF16 || hp->join != VD_JOIN_NONE || pl.kp.K_frames != 1) return false;      // class counts below a compiled shape run padded (weights re-laid out, padding classes masked), like head_kernel
    if (!tfused_supported(pl.C) || hp->T < 1 || hp->frames <= 0 || hp->frames % hp->T || hp->frames / hp->T > 65536) return false;
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        if (!sc.tconv_weight_bf16 || !sc.tconv_scale || !sc.tconv_shift || !sc.tconv_out_nhwc_bf16) return false;
        if (sc.Cin % F_NT || sc.Cin > 1024) return false;
        const long long tiles = (long long)(hp->frames / hp->T) * ceil_div(hp->T * sc.H * sc.W, F_BLOCK_M);
        if (tiles < 2 || tiles > 2 * 65000) return false;      // per-pair item ranges are 16-bit
    }
    return true;
}

static int run_tfused(const VdHeadParams* hp, const HeadPlan& pl, const HeadKernelParams& kp, const void* const* wptr, cudaStream_t stream) {
    const char* e_dbg = getenv("VD_TFUSED_DBG");            // profiling aids (results are garbage): bit 0 = no decode / filter epilogue;
    const char* e_scl = getenv("VD_TFUSED_SCALES");         // bit mask of the scales that get any work
    const int dbg = e_dbg ? atoi(e_dbg) : 0, scl = e_scl ? atoi(e_scl) : 7;
    FusedParams fp;
    FusedMaps maps;
    memset(&fp, 0, sizeof(fp));
    memset(&maps, 0, sizeof(maps));
    fp.B = hp->frames / hp->T; fp.T = hp->T; fp.num_scales = hp->num_scales; fp.slope = 0.1f;
    fp.g = kp.g; fp.c_valid = kp.c_valid; fp.valid_thresh = hp->valid_thresh;
    fp.boxes = kp.boxes; fp.spec_lists = kp.spec_lists; fp.spec_cnt = kp.spec_cnt; fp.spec_tau = kp.spec_tau;
    fp.tile_counter = kp.tile_counter; fp.ws_magic = kp.ws_magic; fp.frames = hp->frames; fp.dbg = dbg; fp.stamps = kp.stamps;
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        FusedScale& q = fp.sc[s];
        q.HW = sc.H * sc.W; q.Cin = sc.Cin; q.rows = hp->T * sc.H * sc.W;
        q.m_tiles = ((scl >> s) & 1) ? ceil_div(q.rows, F_BLOCK_M) : 0; q.n_chunks = sc.Cin / F_NT;
        q.scale = sc.tconv_scale; q.shift = sc.tconv_shift; q.bias = kp.bias[s];
        const uint64_t Cin = (uint64_t)sc.Cin, HW = (uint64_t)sc.H * sc.W;
        const int wstride = sc.tip_window_stride_frames > 0 ? sc.tip_window_stride_frames : hp->T;
        uint64_t dimsX[3] = {Cin, (uint64_t)q.rows, (uint64_t)fp.B};
        uint64_t strX[2] = {Cin * 2, (uint64_t)wstride * HW * Cin * 2};
        uint32_t boxX[3] = {F_BLOCK_K, F_BLOCK_M, 1};
        int rc = encode_tmap_bf16(&maps.x[s], sc.tip_nhwc_bf16, 3, dimsX, strX, boxX);
        if (rc) return rc;
        uint64_t dimsW[3] = {Cin, Cin, 3};
        uint64_t strW[2] = {Cin * 2, Cin * Cin * 2};
        uint32_t boxW[3] = {F_BLOCK_K, F_NT / 2, 1};
        rc = encode_tmap_bf16(&maps.w[s], sc.tconv_weight_bf16, 3, dimsW, strW, boxW);
        if (rc) return rc;
        uint64_t dimsP[2] = {Cin, (uint64_t)3 * (5 + pl.C)};
        uint64_t strP[1] = {Cin * 2};
        uint32_t boxP[2] = {F_BLOCK_K, (uint32_t)(pl.n_pad / 2)};
        rc = encode_tmap_bf16(&maps.wp[s], wptr[s], 2, dimsP, strP, boxP);
        if (rc) return rc;
    }
    // The kernel takes every register of its SMs (168 x 12 warps' worth), so nothing shares them: the previous launch's per-frame NMS
    // kernel runs in the gaps.  Leaving CTA pairs' SMs free for it (VD_TFUSED_SPARE_PAIRS = 1 / 2 / 4) measured 0.817 / 0.832 / 0.855 ms
    // per step against 0.825 with none: no gain, off.
    int spare = 0;
    if (const char* e = getenv("VD_TFUSED_SPARE_PAIRS")) spare = atoi(e);
    int clusters = sm_count() / 2 - (spare > 0 ? spare : 0);
    if (clusters < 1) clusters = 1;
    if (clusters > F_MAX_CLUSTERS) clusters = F_MAX_CLUSTERS;
    tfused_schedule(&fp, clusters);
    return launch_tfused(maps, fp, pl.C, clusters, stream);
}

}  // namespace vd

using namespace vd;

extern "C" size_t vd_sizeof(int which) {
    return which == 0 ? sizeof(VdHeadScale) : (which == 1 ? sizeof(VdHeadParams) : 0);
}

extern "C" size_t vd_head_workspace_bytes(const VdHeadParams* p) {
    HeadPlan pl;
    if (make_plan(p, &pl) != VD_OK) return 0;
    return pl.total;
}

extern "C" size_t vd_head_debug_offset(const VdHeadParams* hp) {
    HeadPlan pl;
    if (make_plan(hp, &pl) != VD_OK) return 0;
    return pl.off_listsA;
}

extern "C" size_t vd_head_stats_offset(const VdHeadParams* hp) {
    HeadPlan pl;
    if (make_plan(hp, &pl) != VD_OK) return 0;
    return pl.off_spec_state;
}

extern "C" int vd_head_launch_count(const VdHeadParams* hp) {
    HeadPlan pl;
    if (make_plan(hp, &pl) != VD_OK) return -1;
    int n = getenv("VD_NO_SPEC") ? 1 + pl.n_pass : 2 + 2 * pl.n_pass;   // head kernel per class window + per-frame NMS kernel (+ the exact fallback pair, idle in the steady state)
    for (int s = 0; s < hp->num_scales; ++s) if (hp->scale[s].tconv_weight_bf16) ++n;
    // fused temporal head: ONE kernel (all scales) instead of the head kernel; the tip cells counted above become the conditional launches
    // of the exact path, behind the kernel that lists the failed windows
    if (tfused_applicable(hp, pl, getenv("VD_NO_SPEC") == nullptr)) ++n;
    if (pl.repack) n += hp->num_scales * pl.n_pass;               // weight re-layout per (scale, window)
    return n;
}

extern "C" int vd_head_fused_tip_plan(const VdHeadParams* hp, int pairs, int* items_out, int* strided_out, int* beg_out) {
    VD_CHECK_ARG(hp && items_out && strided_out && beg_out && pairs >= 1 && pairs <= F_MAX_CLUSTERS, "fused_tip_plan: bad arguments");
    HeadPlan pl;
    int rc = make_plan(hp, &pl);
    if (rc) return rc;
    VD_CHECK_ARG(hp->T >= 1 && hp->frames % hp->T == 0, "fused_tip_plan: frames %d not a multiple of T %d", hp->frames, hp->T);
    FusedParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.B = hp->frames / hp->T; fp.T = hp->T; fp.num_scales = hp->num_scales;
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        fp.sc[s].HW = sc.H * sc.W; fp.sc[s].Cin = sc.Cin; fp.sc[s].rows = hp->T * sc.H * sc.W;
        fp.sc[s].m_tiles = ceil_div(fp.sc[s].rows, F_BLOCK_M); fp.sc[s].n_chunks = sc.Cin / F_NT;
    }
    tfused_schedule(&fp, pairs);
    for (int s = 0; s < VD_MAX_SCALES; ++s) {
        items_out[s] = s < hp->num_scales ? (int)(((long long)fp.B * fp.sc[s].m_tiles + 1) / 2) : 0;
        strided_out[s] = fp.strided[s];
        for (int c = 0; c <= pairs; ++c) beg_out[s * (F_MAX_CLUSTERS + 1) + c] = (int)fp.beg[s][c];
    }
    return VD_OK;
}

extern "C" int vd_head_fused_tip(const VdHeadParams* hp) {
    HeadPlan pl;
    if (make_plan(hp, &pl) != VD_OK) return -1;
    return tfused_applicable(hp, pl, getenv("VD_NO_SPEC") == nullptr) ? 1 : 0;
}

extern "C" int vd_head_forward(const VdHeadParams* hp, float* ids, float* scores, float* bboxes,
                               int32_t* keep_rows_or_null, void* workspace, size_t workspace_bytes, void* stream_) {
    return vd_head_forward_stages(hp, ids, scores, bboxes, keep_rows_or_null, workspace, workspace_bytes, stream_,
                                  VD_STAGE_ALL);
}

extern "C" int vd_head_forward_stages(const VdHeadParams* hp, float* ids, float* scores, float* bboxes,
                                      int32_t* keep_rows_or_null, void* workspace, size_t workspace_bytes,
                                      void* stream_, int stage_mask) {
    cudaStream_t stream = (cudaStream_t)stream_;
    HeadPlan pl;
    int rc = make_plan(hp, &pl);
    if (rc) return rc;
    VD_CHECK_ARG(hp->frames == 0 || (ids && scores && bboxes), "head_forward: null output");
    VD_CHECK_ARG(hp->nms_thresh > 0.f && hp->nms_thresh < 1.f,
                 "head_forward: nms_thresh %g disables NMS (yolo3.py:525); use vd_head_detections for the raw rows", hp->nms_thresh);
    VD_CHECK_ARG(hp->post_nms > 0, "head_forward: post_nms must be > 0 (use vd_head_detections + vd_box_nms for the unsliced output)");
    const long long rows = pl.kp.g.row_base[hp->num_scales];
    long long k64 = (hp->nms_topk > 0 && hp->nms_topk < rows) ? hp->nms_topk : rows;
    if (k64 > VD_MAX_TOPK) return set_error(VD_ERR_UNSUPPORTED, "head_forward: nms_topk %lld > VD_MAX_TOPK %d", k64, VD_MAX_TOPK);
    if (hp->frames > 0 && (!workspace || workspace_bytes < pl.total)) return set_error(VD_ERR_WORKSPACE, "head_forward: workspace %zu < required %zu", workspace_bytes, pl.total);
    VD_CHECK_ARG(((uintptr_t)bboxes & 15) == 0, "head_forward: bboxes must be 16-byte aligned");
    if (hp->frames == 0) return VD_OK;
    const int k = (int)k64;
    unsigned char* ws = (unsigned char*)workspace;
    HeadKernelParams& kp = pl.kp;
    kp.hints = (unsigned long long*)(ws + pl.off_hints);
    kp.hist = (uint32_t*)(ws + pl.off_hist);
    kp.tile_counter = (unsigned int*)(ws + pl.off_ctr);
    kp.ws_magic = kWsMagic ^ (unsigned int)(pl.total * 2654435761u) ^ ((unsigned int)hp->frames << 20) ^ ((unsigned int)kp.tiles_per_frame << 8);
    kp.boxes = (float4*)(ws + pl.off_boxes);
    kp.lists = (uint64_t*)(ws + pl.off_lists0);
    kp.counts = (uint32_t*)(ws + pl.off_counts0);
    kp.counts_hi = (uint32_t*)(ws + pl.off_counts_hi);
    kp.hint_hi = (uint32_t*)(ws + pl.off_hint_hi);
    kp.coarse = (uint32_t*)(ws + pl.off_coarse);
    kp.spec_lists = (uint64_t*)(ws + pl.off_spec_lists);
    kp.spec_cnt = (uint32_t*)(ws + pl.off_spec_cnt);
    kp.spec_state = (uint32_t*)(ws + pl.off_spec_state);
    kp.spec_tau = (uint32_t*)(ws + pl.off_spec_tau);
    uint32_t* failed = (uint32_t*)(ws + pl.off_failed);
    const bool spec = getenv("VD_NO_SPEC") == nullptr;   // speculative frame-level threshold with the exact path as fallback
    kp.valid_thresh = hp->valid_thresh; kp.k = k; kp.cap = (k <= 448) ? 640 : kListCap;
    { const char* e = getenv("VD_HEAD_PREFETCH"); kp.prefetch = e ? atoi(e) : 0; }
    kp.stamps = (getenv("VD_DEBUG_HEAD_STAMPS") || getenv("VD_TFUSED_STAMPS")) ? (long long*)(ws + pl.off_listsA) : nullptr;   // profiling aid (unused merge area)
    if (const char* e = getenv("VD_DEBUG_SKIP_EPILOGUE")) kp.dbg = atoi(e);     // profiling aid: results are garbage

    const bool fused_tip = tfused_applicable(hp, pl, spec);      // tip cell + head in one kernel per scale (tfused.cuh): the TCONV stage is empty
    for (int s = 0; s < hp->num_scales; ++s) {      // optional temporal tip cell in front (layers.py:82-89)
        const VdHeadScale& sc = hp->scale[s];
        if (sc.tconv_weight_bf16 && (stage_mask & VD_STAGE_TCONV) && !fused_tip) {
            VD_CHECK_ARG(sc.tconv_out_nhwc_bf16 && sc.tconv_scale && sc.tconv_shift, "head_forward: scale %d temporal cell needs out/scale/shift", s);
            VD_CHECK_ARG(hp->T >= 1 && hp->frames % hp->T == 0, "head_forward: frames %d not a multiple of T %d", hp->frames, hp->T);
            rc = vd_temporal_conv_ex(sc.tip_nhwc_bf16, sc.tconv_out_nhwc_bf16, hp->frames / hp->T, hp->T, sc.H, sc.W, sc.Cin,
                                     sc.tconv_weight_bf16, sc.tconv_scale, sc.tconv_shift, 0.1f, VD_PREC_BF16, sc.tip_window_stride_frames, stream_);
            if (rc) return rc;
        }
    }
    const void* wptr[VD_MAX_SCALES];
    HeadMaps maps;
    if (stage_mask & VD_STAGE_HEAD) {
        // no memset: the previous call's NMS kernels left the counters / histograms zeroed (layout marker); a workspace in
        // any other state is detected on the device and handled exactly
        rc = repack_windows(hp, pl, ws, stream);
        if (rc) return rc;
        const bool pair_head = !fused_tip && hpair_applicable(hp, pl, spec);      // wide heads: the speculative kernel on CTA pairs (hpair.cuh)
        if (fused_tip || pair_head) {
            window_params(hp, pl, ws, 0, &kp, wptr);
            rc = fused_tip ? run_tfused(hp, pl, kp, wptr, stream) : run_hpair(hp, pl, kp, wptr, stream);
            if (rc) return rc;
        }
        for (int ps = 0; ps < pl.n_pass && !fused_tip && !pair_head; ++ps) {       // one launch per class window (1 unless num_class > 80), all appending to the frames' lists
            window_params(hp, pl, ws, ps, &kp, wptr);
            rc = make_maps(hp, pl, &maps, wptr);
            if (rc) return rc;
            rc = spec ? launch_head<EPI_SPEC>(maps, kp, pl.C, stream) : launch_head<EPI_FILTER>(maps, kp, pl.C, stream);
            if (rc) return rc;
        }
    }
    if (!(stage_mask & VD_STAGE_NMS)) return VD_OK;

    NmsParams P;
    P.overlap_thresh = hp->nms_thresh; P.k = k; P.sortn = nms_sortn(k); P.class_aware = 1;
    P.max_out = hp->post_nms;
    P.dbg = (getenv("VD_DEBUG_NMS_STAMPS") || getenv("VD_DEBUG_HEAD_STAMPS")) ? (long long*)(ws + pl.off_listsA) : nullptr;   // profiling aid (unused merge area)
    FusedSource src{kp.g, kp.boxes};
    FusedSink sink;
    sink.ids = ids; sink.scores = scores; sink.bboxes = bboxes; sink.keep = keep_rows_or_null; sink.post = hp->post_nms;
    VD_CHECK_ARG(hp->n_mirrors >= 0 && hp->n_mirrors <= VD_MAX_MIRRORS, "head_forward: n_mirrors %d", hp->n_mirrors);
    sink.n_mirror = hp->n_mirrors;
    for (int i = 0; i < VD_MAX_MIRRORS; ++i) {
        sink.mirror[i] = i < hp->n_mirrors ? hp->mirror_delta[i] : 0;
        VD_CHECK_ARG((sink.mirror[i] & 15) == 0, "head_forward: mirror_delta[%d] must be a multiple of 16 bytes", i);
    }
    {
        const bool carve = !getenv("VD_DEBUG_NO_CARVEOUT") || atoi(getenv("VD_DEBUG_NO_CARVEOUT")) == 2;
        rc = configure_kernel((const void*)nms_final_hist_kernel, (int)nms_hist_smem(VD_MAX_TOPK, VD_MAX_TOPK), carve);
        if (rc) return rc;
        rc = configure_kernel((const void*)nms_spec_kernel, (int)nms_spec_smem(VD_MAX_TOPK, VD_MAX_TOPK), carve);
        if (rc) return rc;
    }
    if (!spec) {
        nms_final_hist_kernel<<<hp->frames, kNmsThreads, nms_hist_smem(k, hp->post_nms), stream>>>(
            kp.lists, kp.counts, kp.counts_hi, kp.hint_hi, kp.coarse, kp.tiles_per_frame * pl.n_pass, kp.hist, kp.tile_counter, kp.ws_magic,
            nullptr, nullptr, nullptr, P, src, sink);
        VD_LAUNCH_CHECK();
        return VD_OK;
    }
    // 1. frames whose speculative list is provably complete are finished; the others are queued in `failed`
    nms_spec_kernel<<<hp->frames, kNmsThreads, nms_spec_smem(k, hp->post_nms), stream>>>(
        kp.spec_lists, kp.spec_cnt, kp.spec_state, kp.spec_tau, failed, kp.tile_counter, kp.ws_magic, hp->valid_thresh, P, src, sink);
    VD_LAUNCH_CHECK();
    // 2. exact path over the queued frames (both kernels return at once when the queue is empty -- the steady state)
    if (fused_tip) {
        // the fused kernel kept the tip on chip: the exact path reads it from memory, so the tip cells run first -- only if a frame failed,
        // and only over the windows of the failed frames
        uint32_t* winlist = (uint32_t*)(ws + pl.off_winlist);
        failed_windows_kernel<<<1, 1024, 0, stream>>>(failed, kp.spec_state + 2, hp->T, hp->frames / hp->T, winlist);
        VD_LAUNCH_CHECK();
        for (int s = 0; s < hp->num_scales; ++s) {
            const VdHeadScale& sc = hp->scale[s];
            rc = temporal_conv_impl(sc.tip_nhwc_bf16, sc.tconv_out_nhwc_bf16, hp->frames / hp->T, hp->T, sc.H, sc.W, sc.Cin, sc.tconv_weight_bf16,
                                    sc.tconv_scale, sc.tconv_shift, 0.1f, VD_PREC_BF16, sc.tip_window_stride_frames, winlist, winlist + 64, stream_);
            if (rc) return rc;
        }
    }
    for (int ps = 0; ps < pl.n_pass; ++ps) {
        HeadKernelParams kf = kp;
        window_params(hp, pl, ws, ps, &kf, wptr);
        rc = make_maps(hp, pl, &maps, wptr);
        if (rc) return rc;
        kf.frame_list = failed; kf.frame_count = kp.spec_state + 2;
        kf.total_tiles = kp.tiles_per_frame * hp->frames;              // grid sizing only: the device reads the real count
        rc = launch_head<EPI_FILTER>(maps, kf, pl.C, stream);
        if (rc) return rc;
    }
    nms_final_hist_kernel<<<hp->frames, kNmsThreads, nms_hist_smem(k, hp->post_nms), stream>>>(
        kp.lists, kp.counts, kp.counts_hi, kp.hint_hi, kp.coarse, kp.tiles_per_frame * pl.n_pass, kp.hist, kp.tile_counter, kp.ws_magic,
        failed, kp.spec_state, kp.spec_tau, P, src, sink);
    VD_LAUNCH_CHECK();
    return VD_OK;
}

extern "C" int vd_head_detections(const VdHeadParams* hp, float* det, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    HeadPlan pl;
    int rc = make_plan(hp, &pl);
    if (rc) return rc;
    VD_CHECK_ARG(hp->frames == 0 || det, "head_detections: null output");
    VD_CHECK_ARG(((uintptr_t)det & 7) == 0, "head_detections: det must be 8-byte aligned");
    if (hp->frames == 0) return VD_OK;
    HeadKernelParams& kp = pl.kp;
    kp.det = det; kp.det_rows_total = kp.g.row_base[hp->num_scales];
    for (int s = 0; s < hp->num_scales; ++s) {
        const VdHeadScale& sc = hp->scale[s];
        if (sc.tconv_weight_bf16) {
            VD_CHECK_ARG(sc.tconv_out_nhwc_bf16 && sc.tconv_scale && sc.tconv_shift, "head_detections: scale %d temporal cell needs out/scale/shift", s);
            VD_CHECK_ARG(hp->T >= 1 && hp->frames % hp->T == 0, "head_detections: frames %d not a multiple of T %d", hp->frames, hp->T);
            rc = vd_temporal_conv_ex(sc.tip_nhwc_bf16, sc.tconv_out_nhwc_bf16, hp->frames / hp->T, hp->T, sc.H, sc.W, sc.Cin,
                                     sc.tconv_weight_bf16, sc.tconv_scale, sc.tconv_shift, 0.1f, VD_PREC_BF16, sc.tip_window_stride_frames, stream_);
            if (rc) return rc;
        }
    }
    if (pl.repack && (!workspace || workspace_bytes < pl.total))
        return set_error(VD_ERR_WORKSPACE, "head_detections: workspace %zu < required %zu (class windows need the weight re-layout area)", workspace_bytes, pl.total);
    unsigned char* ws = (unsigned char*)workspace;
    rc = repack_windows(hp, pl, ws, stream);
    if (rc) return rc;
    const void* wptr[VD_MAX_SCALES];
    HeadMaps maps;
    for (int ps = 0; ps < pl.n_pass; ++ps) {
        window_params(hp, pl, ws, ps, &kp, wptr);
        rc = make_maps(hp, pl, &maps, wptr);
        if (rc) return rc;
        rc = launch_head<EPI_DET>(maps, kp, pl.C, stream);
        if (rc) return rc;
    }
    return VD_OK;
}

extern "C" int vd_pred_conv(const void* x, int B, int H, int W, int Cin, int K_frames, int join,
                            const void* weight, const float* bias, int N, float* pred, void* stream_) {
    return vd_pred_conv_ex(x, B, H, W, Cin, K_frames, join, VD_PREC_BF16, weight, bias, N, pred, stream_);
}

extern "C" int vd_pred_conv_ex(const void* x, int B, int H, int W, int Cin, int K_frames, int join, int precision,
                               const void* weight, const float* bias, int N, float* pred, void* stream_) {
    VD_CHECK_ARG(weight && (B == 0 || (x && pred)), "pred_conv: null pointer");
    VD_CHECK_ARG(precision == VD_PREC_BF16 || precision == VD_PREC_FP32_SPLIT || precision == VD_PREC_BF16X2, "pred_conv: precision %d", precision);
    const bool split = precision != VD_PREC_BF16;
    const int planes = precision == VD_PREC_FP32_SPLIT ? 3 : 2;
    if (split && join != VD_JOIN_NONE) return set_error(VD_ERR_UNSUPPORTED, "pred_conv: the fp32-parity modes are not combinable with a late cat join");
    VD_CHECK_ARG(B >= 0 && B <= 65535 && H > 0 && W > 0 && N > 0, "pred_conv: bad shape");
    VD_CHECK_ARG(Cin > 0 && Cin % BLOCK_K == 0, "pred_conv: Cin %d must be a multiple of %d", Cin, BLOCK_K);
    VD_CHECK_ARG(join == VD_JOIN_NONE || join == VD_JOIN_CAT, "pred_conv: join %d must be pre-reduced (vd_temporal_pool)", join);
    VD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)weight & 15) == 0, "pred_conv: tensors must be 16-byte aligned");
    if (B == 0) return VD_OK;
    const int K = split ? planes : ((join == VD_JOIN_CAT) ? K_frames : 1);      // operand planes / joined frames in memory
    VD_CHECK_ARG(K >= 1, "pred_conv: K_frames %d", K_frames);
    const uint64_t HW = (uint64_t)H * W;
    for (int n0 = 0; n0 < N; n0 += 256) {            // output-channel slices of <= 256 rows of W
        const int nn = (N - n0) < 256 ? (N - n0) : 256;
        HeadKernelParams kp; memset(&kp, 0, sizeof(kp));
        kp.g.num_scales = 1; kp.g.num_class = 1; kp.g.A = 3;
        kp.g.H[0] = H; kp.g.W[0] = W; kp.g.HW[0] = (int)HW; kp.g.stride[0] = 1.f;
        kp.frames = B; kp.K_frames = K; kp.cin[0] = Cin;
        if (split) split_products(precision, &kp);
        kp.pb[0] = ceil_div((int)HW, BLOCK_M);
        kp.tile_start[0] = 0; kp.tile_start[1] = kp.pb[0] * B; kp.order[0] = 0;
        kp.tiles_per_frame = kp.pb[0]; kp.total_tiles = kp.pb[0] * B;
        kp.n_pad = (nn + 15) / 16 * 16; kp.n_valid = nn;
        kp.bias[0] = bias ? bias + n0 : nullptr;
        kp.pred[0] = pred + (size_t)n0 * HW; kp.pred_frame_stride = (long long)N * (long long)HW;
        HeadMaps maps;
        uint64_t dimsA[4] = {(uint64_t)Cin, HW, (uint64_t)K, (uint64_t)B};
        uint64_t strA[3] = {(uint64_t)Cin * 2, HW * Cin * 2, (uint64_t)K * HW * Cin * 2};
        if (split) { strA[1] = (uint64_t)B * HW * Cin * 2; strA[2] = HW * Cin * 2; }      // plane-major (planes, B, H, W, Cin)
        uint32_t boxA[4] = {BLOCK_K, BLOCK_M, 1, 1};
        int rc = encode_tmap_bf16(&maps.a[0], x, 4, dimsA, strA, boxA);
        if (rc) return rc;
        uint64_t dimsW[2] = {(uint64_t)Cin * K, (uint64_t)nn};
        uint64_t strW[1] = {(uint64_t)Cin * K * 2};
        uint32_t boxW[2] = {BLOCK_K, (uint32_t)kp.n_pad};
        rc = encode_tmap_bf16(&maps.w[0], (const unsigned char*)weight + (size_t)n0 * Cin * K * 2, 2, dimsW, strW, boxW);
        if (rc) return rc;
        maps.a[1] = maps.a[2] = maps.a[0]; maps.w[1] = maps.w[2] = maps.w[0];
        rc = launch_pred(maps, kp, (cudaStream_t)stream_);
        if (rc) return rc;
    }
    return VD_OK;
}
