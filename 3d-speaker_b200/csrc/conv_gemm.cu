// 1x1-convolution GEMM for sm_100a: both operands by TMA, BN-ReLU prologue applied in shared
// memory, tcgen05.mma with TMEM accumulators.  Serves the layers that hold half of CAM++'s
// FLOPs: the D-TDNN bottleneck 1x1s over the growing concat buffer (layers.py:140-141), the
// transit layers (layers.py:193-196) and every other stride-1 1x1 conv of the embedding nets.
//
//   D[m, n] = sum_k f(A[m, k]) * W[n, k],   f = identity or relu(x*scale_k + shift_k)
//
// A is the channels-last activation matrix [M, ld] (a channel window of a wider buffer), W the
// packed [Cout, K] weights; both are loaded as 128-byte-swizzled K-major tiles by
// cp.async.bulk.tensor.2d (out-of-range rows / channels are zero-filled by the TMA unit), so
// the loads need no registers and run a full stage ring (5-6 x 32 KB) ahead of the tensor core.
// The pre-activation BatchNorm+ReLU of the dense block cannot be folded into the producer of
// the concat buffer (every layer applies its own BN to all earlier channels), so four
// "transform" warps apply it to the landed A tile with packed bf16x2 math (HFMA2.BF16 + HMNMX2).
// For N <= 128 the transformed tile goes to TENSOR MEMORY (thread = tile row = TMEM lane,
// tcgen05.st) and the MMA takes A from there: the kernel is bound by shared-memory bandwidth
// (TMA writes + transform reads + MMA operand reads), and this drops the write-back and the MMA's
// A read, 96 -> 64 KB of shared-memory traffic per 128x128x64 stage.  For N = 256 the accumulators
// fill TMEM and the tile is rewritten in place in shared memory instead.
//
// Convs that are not 1x1 stride 1 (the 3x3 convs of ERes2NetV2 on 96-208 channels, strided 1x1 shortcuts, the CAM++ tdnn
// layer) run through the same kernel with the A tile fetched by TMA IM2COL loads: a tensor map over the [B, H, W, C]
// activations whose pixel box is the set of filter base positions; one cp.async.bulk.tensor.4d...im2col per (filter
// tap, 64-channel chunk) delivers the 128 consecutive output pixels of the tile (walking W, then H, then B with the
// conv stride, out-of-image taps zero-filled) as the same 128B-swizzled K-major tile.  The K loop is taps x chunks;
// a chunk that runs past Cin reads zeros (the map's channel axis ends at the conv's last input channel).  The
// gather kernel this replaces was bound by its 16-byte loads at ~3 TB/s of L2 traffic (conv_tc2.cu).
//
// Warp roles (320 threads, one persistent CTA per SM): 0 TMA producer, 1 TMEM alloc + MMA issue,
// 2-5 transform, 6-9 epilogue (folded BN, residual, activation, CAM gate; double-buffered TMEM
// accumulator so the epilogue of tile i overlaps the mainloop of tile i+1).
#include <cstdlib>
#include <cuda.h>

#include <algorithm>
#include <map>
#include <mutex>

#include "ops.cuh"
#include "tc.cuh"
#include "tmap.cuh"

namespace spk {
namespace {

using namespace tc;        // mbarrier / TMEM / tcgen05 wrappers shared by all tensor-core kernels
using bf16 = __nv_bfloat16;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kThreads = 480;      // 0 TMA, 1 MMA, 2-5 transform A, 6-9 epilogue, 10-13 transform B, 14 TMA store / residual load
constexpr int kXformThreads = 128;
constexpr int kEpilogueThreads = 128;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D (+)= A * B^T with A read from tensor memory (lane = row, one 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_bf16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap *map, int c, int w, int h, int n, uint16_t off_w,
                                                   uint16_t off_h, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// SW128 K-major descriptor from 32-bit halves (stepping K by 16 elements = +2 on the low half)
__device__ __forceinline__ uint32_t sw128_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kSw128Hi = (1024u >> 4) | (1u << 14) | (2u << 29);
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// debug aid: per-stage role timestamps of CTA 0 (SPK_GEMM_DBG=1), read back by spk_debug_gemm_timeline
__device__ long long g_gemm_ts[256 * 8];
#define GEMM_TS(idx, slot) do { if (dbg && blockIdx.x == 0 && (idx) < 256 && (threadIdx.x & 31) == 0) g_gemm_ts[(idx) * 8 + (slot)] = clock64(); } while (0)

template <int BLOCK_N, bool TEPI = false> struct Cfg {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    // TEPI (TMA epilogue): output tiles leave in UNITS of 128 rows x <= 128 columns through two staging buffers, so
    // the ring is shorter.  A unit is one or two 128B-swizzled blocks of 128 rows x 64 columns.
    static constexpr int kUnitBlocks = BLOCK_N >= 128 ? 2 : 1;
    static constexpr int kUnitBytes = TEPI ? kUnitBlocks * BLOCK_M * 128 : 0;
    static constexpr int kUnitsPerTile = (BLOCK_N + 127) / 128;
    static constexpr int kStgBytes = 2 * kUnitBytes;
    static constexpr int kStages = TEPI ? ((BLOCK_N >= 256) ? 3 : (BLOCK_N >= 128 ? 4 : 6)) : ((BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8));
    static constexpr bool kATmem = BLOCK_N <= 128;          // room for the A ring next to the two accumulators
    static constexpr int kAColsPerStage = BLOCK_K / 2;       // two bf16 per 32-bit TMEM column
    static constexpr int kTmemNeed = 2 * BLOCK_N + (kATmem ? kStages * kAColsPerStage : 0);
    static constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStgBytes + 1024 + 512;
    // the two transform groups of the tensor-memory prologue take alternate stage USES; with an even ring every
    // stage (and its barriers) belongs to one group.  With an odd ring consecutive uses of a stage would alternate
    // between the groups and a group could wait on a barrier two phases ahead, which parity waits cannot tell apart.
    static_assert(!kATmem || kStages % 2 == 0, "tensor-memory prologue needs an even stage ring");
};

template <int BLOCK_N, typename TOut, typename TRes, bool TEPI>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const ConvArgs a, const bf16 *__restrict__ pro_scale_bf, const bf16 *__restrict__ pro_shift_bf,
                 int n_tiles_n, long long n_tiles, const __grid_constant__ CUtensorMap amap,
                 const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap ymap,
                 const __grid_constant__ CUtensorMap rmap, int dbg, int im2col) {
    using C = Cfg<BLOCK_N, TEPI>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t s_stg = base + C::kStages * C::kStageBytes;       // TEPI: output staging tile
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (base - raw) + C::kStages * C::kStageBytes + C::kStgBytes);
    const uint32_t bar0 = smem_u32(bars);
    // [0,S) landed (TMA tx)  [S,2S) ready (transformed)  [2S,3S) empty  then accum full/empty x2
    auto land_bar = [&](int s) { return bar0 + 8u * s; };
    auto ready_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + s); };
    auto accf_bar = [&](int b) { return bar0 + 8u * (3 * C::kStages + b); };
    auto acce_bar = [&](int b) { return bar0 + 8u * (3 * C::kStages + 2 + b); };
    auto rfull_bar = [&](uint32_t b) { return bar0 + 8u * (3 * C::kStages + 4 + b); };     // TEPI: residual unit landed in staging buffer b
    auto sfree_bar = [&](uint32_t b) { return bar0 + 8u * (3 * C::kStages + 6 + b); };     // TEPI: the TMA store has read staging buffer b
    auto gfull_bar = [&](uint32_t b) { return bar0 + 8u * (3 * C::kStages + 8 + b); };     // TEPI: staging buffer b holds a finished unit
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bars + 3 * C::kStages + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_pro = pro_scale_bf != nullptr;
    // Without the tensor-memory prologue warps 10-13 have nothing to transform: they form a second epilogue group.
    // One unit per tile (N <= 128): the groups take alternate tiles (= alternate accumulators and staging buffers);
    // two units per tile (N = 256): group g takes the tile's column half g.  The epilogue - one TMEM round trip per 32
    // columns and row - was the per-tile critical path of every GEMM with a short K loop.
    const bool two_groups = !(has_pro && C::kATmem);
    constexpr bool kSplitCols = C::kUnitsPerTile == 2;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(land_bar(s), 1);
            mbar_init(ready_bar(s), kXformThreads);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accf_bar(b), 1);
            mbar_init(acce_bar(b), kEpilogueThreads * ((kSplitCols && two_groups) ? 2 : 1));
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(rfull_bar(b), 1);
            mbar_init(sfree_bar(b), 1);
            mbar_init(gfull_bar(b), kEpilogueThreads);
        }
        if (TEPI) {
            prefetch_tmap(&ymap);
            prefetch_tmap(&rmap);
        }
        fence_barrier_init();
        prefetch_tmap(&amap);
        prefetch_tmap(&wmap);
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), C::kTmemCols);
    // prologue scale/shift -> shared memory (zero beyond K: relu(0*x+0) = 0), read back as broadcast loads
    const int nk_pre = (a.K + BLOCK_K - 1) / BLOCK_K;
    const uint32_t s_pro = base + C::kStages * C::kStageBytes + C::kStgBytes + 512u, pro_bytes = (uint32_t)nk_pre * 128u;
    if (has_pro && C::kATmem) {
        for (int idx = threadIdx.x; idx < nk_pre * 8; idx += kThreads) {
            const int c = idx * 8;
            uint4 s4 = make_uint4(0u, 0u, 0u, 0u), h4 = s4;
            if (c < a.K) {
                s4 = ldg16(pro_scale_bf + c);
                h4 = ldg16(pro_shift_bf + c);
            }
            sts16(s_pro + (uint32_t)idx * 16u, s4);
            sts16(s_pro + pro_bytes + (uint32_t)idx * 16u, h4);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int chunks = (a.Cin + BLOCK_K - 1) / BLOCK_K;                      // im2col: 64-channel chunks per filter tap
    const int nk = im2col ? a.KH * a.KW * chunks : (a.K + BLOCK_K - 1) / BLOCK_K;
    pdl_wait();          // everything above read only parameters; the activations come from the previous kernel

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int stage = 0, sidx = 0;
            uint32_t phase = 0;
            const uint64_t pol_first = l2_policy_evict_first();
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const long long mt = tile / n_tiles_n;
                const int nt = (int)(tile - mt * n_tiles_n);
                // im2col: base position (filter tap 0, 0) of the tile's first output pixel
                int bw = 0, bh = 0, bn = 0, kh = 0, kw = 0, ch = 0;
                if (im2col) {
                    const long long m0 = mt * BLOCK_M;
                    const int hw = a.Ho * a.Wo;
                    bn = (int)(m0 / hw);
                    const int r = (int)(m0 - (long long)bn * hw);
                    const int p = r / a.Wo;
                    bh = p * a.sh - a.ph;
                    bw = (r - p * a.Wo) * a.sw - a.pw;
                }
                for (int kc = 0; kc < nk; ++kc, ++sidx) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    GEMM_TS(sidx, 0);
                    const uint32_t sa = base + stage * C::kStageBytes;
                    mbar_arrive_expect_tx(land_bar(stage), C::kStageBytes);
                    if (im2col) {
                        tma_load_im2col_4d(sa, &amap, a.in_choff + ch * BLOCK_K, bw, bh, bn, (uint16_t)(kw * a.dw), (uint16_t)(kh * a.dh),
                                           land_bar(stage));
                        tma_load_2d(sa + C::kABytes, &wmap, (kh * a.KW + kw) * a.Cin + ch * BLOCK_K, nt * BLOCK_N, land_bar(stage));
                        if (++ch == chunks) { ch = 0; if (++kw == a.KW) { kw = 0; ++kh; } }
                    } else {
                        if (a.l2_flags & 1) tma_load_2d_hint(sa, &amap, kc * BLOCK_K, (int)(mt * BLOCK_M), land_bar(stage), pol_first);
                        else tma_load_2d(sa, &amap, kc * BLOCK_K, (int)(mt * BLOCK_M), land_bar(stage));
                        tma_load_2d(sa + C::kABytes, &wmap, kc * BLOCK_K, nt * BLOCK_N, land_bar(stage));
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        constexpr uint32_t idesc = make_idesc(BLOCK_N);
        int stage = 0, sidx = 0;
        uint32_t phase = 0, it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, acc_phase = (it >> 1) & 1u;
            mbar_wait(acce_bar(buf), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
            for (int kc = 0; kc < nk; ++kc, ++sidx) {
                GEMM_TS(sidx, 3);
                mbar_wait(has_pro ? ready_bar(stage) : land_bar(stage), phase);
                tc_fence_after();
                GEMM_TS(sidx, 4);
                if (elect_one()) {
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint32_t lo_a = sw128_lo(sa), lo_b = sw128_lo(sa + C::kABytes);
                    if (C::kATmem && has_pro) {          // A = the transformed tile in tensor memory
                        const uint32_t ta = tmem_base + 2 * BLOCK_N + stage * C::kAColsPerStage;
                        if (kc == 0) umma_bf16_ta(d_tmem, ta, desc64(lo_b, kSw128Hi), idesc, 0u);
                        else umma_bf16_ta(d_tmem, ta, desc64(lo_b, kSw128Hi), idesc, 1u);
#pragma unroll
                        for (int kk = 1; kk < BLOCK_K / UMMA_K; ++kk)
                            umma_bf16_ta(d_tmem, ta + kk * (UMMA_K / 2), desc64(lo_b + 2u * kk, kSw128Hi), idesc, 1u);
                    } else {
                        if (kc == 0) umma_bf16(d_tmem, desc64(lo_a, kSw128Hi), desc64(lo_b, kSw128Hi), idesc, 0u);
                        else umma_bf16(d_tmem, desc64(lo_a, kSw128Hi), desc64(lo_b, kSw128Hi), idesc, 1u);
#pragma unroll
                        for (int kk = 1; kk < BLOCK_K / UMMA_K; ++kk)
                            umma_bf16(d_tmem, desc64(lo_a + 2u * kk, kSw128Hi), desc64(lo_b + 2u * kk, kSw128Hi), idesc, 1u);
                    }
                    umma_commit(empty_bar(stage));
                    if (kc == nk - 1) umma_commit(accf_bar(buf));
                }
                __syncwarp();
                GEMM_TS(sidx, 5);
                if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp < 6 || (warp >= 10 && warp < 14 && !two_groups)) {
        // =========================== transform (BN-ReLU prologue) ===========================
        // TMEM path: two groups of four warps (2-5 and 10-13) take alternate stages, so the latency of one stage's
        // load -> math -> tcgen05.st chain overlaps the next stage's
        if (has_pro && C::kATmem) {
            // thread = tile row = TMEM lane: read the row's eight 16-byte chunks (swizzled), transform, store the
            // 64 bf16 as 32 columns of this lane in the stage's TMEM slot
            const int q = warp & 3, r = q * 32 + lane;
            const int grp = warp >= 10 ? 1 : 0;
            const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128), x = (uint32_t)(r & 7);
            const bool relu = a.pro_relu != 0;
            const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
            const long long n_stage_uses = my_tiles * nk;
            for (long long sidx = grp; sidx < n_stage_uses; sidx += 2) {
                const int stage = (int)(sidx % C::kStages);
                const uint32_t phase = (uint32_t)((sidx / C::kStages) & 1);
                const int kc = (int)(sidx % nk);
                mbar_wait(land_bar(stage), phase);
                if (lane == 0 && q == 2) GEMM_TS(sidx, 1);
                const uint32_t sa = base + stage * C::kStageBytes + row_off;
                const uint32_t sp = s_pro + (uint32_t)kc * 128u;          // this stage's 64 scales, 64 shifts at +pro_bytes
                uint32_t out[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 s4 = lds16(sp + j * 16), h4 = lds16(sp + pro_bytes + j * 16);
                    const uint4 v = lds16(sa + (((uint32_t)j ^ x) << 4));
                    out[4 * j] = bnrelu2(v.x, s4.x, h4.x, relu);
                    out[4 * j + 1] = bnrelu2(v.y, s4.y, h4.y, relu);
                    out[4 * j + 2] = bnrelu2(v.z, s4.z, h4.z, relu);
                    out[4 * j + 3] = bnrelu2(v.w, s4.w, h4.w, relu);
                }
                tmem_st32(tmem_base + 2 * BLOCK_N + stage * C::kAColsPerStage + ((uint32_t)(q * 32) << 16), out);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(ready_bar(stage));
                if (lane == 0 && q == 2) GEMM_TS(sidx, 2);
            }
        } else if (has_pro && warp < 6) {
            const int t = threadIdx.x - 64;            // 0..127
            const int j = t & 7, r0 = t >> 3;          // 16-byte column j, rows r0 + 16*i
            const uint32_t sw_off = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
            const bool relu = a.pro_relu != 0;
            int stage = 0, sidx = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < nk; ++kc, ++sidx) {
                    // scale/shift of this thread's 8 channels (zero beyond K: relu(0*x+0) = 0)
                    const int c = kc * BLOCK_K + j * 8;
                    uint4 s4 = make_uint4(0u, 0u, 0u, 0u), h4 = s4;
                    if (c < a.K) {
                        s4 = ldg16(pro_scale_bf + c);
                        h4 = ldg16(pro_shift_bf + c);
                    }
                    mbar_wait(land_bar(stage), phase);
                    const uint32_t sa = base + stage * C::kStageBytes + sw_off;
                    uint4 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = lds16(sa + i * 2048);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i].x = bnrelu2(v[i].x, s4.x, h4.x, relu);
                        v[i].y = bnrelu2(v[i].y, s4.y, h4.y, relu);
                        v[i].z = bnrelu2(v[i].z, s4.z, h4.z, relu);
                        v[i].w = bnrelu2(v[i].w, s4.w, h4.w, relu);
                        sts16(sa + i * 2048, v[i]);
                    }
                    fence_proxy_async();
                    mbar_arrive(ready_bar(stage));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 14) {
        // =========================== TMA store / residual load (TMA epilogue only) ===========================
        if (TEPI && elect_one()) {
            // units in issue order: tile by tile, column half by column half; staging buffer = column half (two units
            // per tile) or tile parity (one unit per tile); barrier phase = that buffer's use count
            struct Unit { long long tile; uint32_t it; int h; };
            auto units_of = [&](long long tile) { return min(C::kUnitsPerTile, (a.Cout - (int)(tile % n_tiles_n) * BLOCK_N + 127) / 128); };
            auto next = [&](Unit &u) {
                if (u.h + 1 < units_of(u.tile)) ++u.h;
                else { u.h = 0; u.tile += gridDim.x; ++u.it; }
            };
            auto sb_of = [&](const Unit &u) { return kSplitCols ? (uint32_t)u.h : (u.it & 1u); };
            auto next_same = [&](Unit &u) {
                const uint32_t sb = sb_of(u);
                do next(u); while (u.tile < n_tiles && sb_of(u) != sb);
            };
            auto coords = [&](const Unit &u, int &row0, int &col0, int &nblk) {
                const long long mt = u.tile / n_tiles_n;
                const int nt = (int)(u.tile - mt * n_tiles_n);
                row0 = (int)(mt * BLOCK_M);
                col0 = nt * BLOCK_N + u.h * 128;
                nblk = min(C::kUnitBlocks, (a.Cout - col0 + 63) / 64);
            };
            const bool has_res = a.res != nullptr;
            auto load_res = [&](const Unit &u, uint32_t sb) {
                int row0, col0, nblk;
                coords(u, row0, col0, nblk);
                mbar_arrive_expect_tx(rfull_bar(sb), (uint32_t)nblk * (BLOCK_M * 128u));
                for (int j = 0; j < nblk; ++j)
                    tma_load_2d(s_stg + sb * C::kUnitBytes + (uint32_t)j * (BLOCK_M * 128u), &rmap, col0 + 64 * j, row0, rfull_bar(sb));
            };
            const uint64_t pol_last = l2_policy_evict_last();
            Unit cur{(long long)blockIdx.x, 0u, 0}, pre[2];
            if (has_res)
                for (uint32_t sb = 0; sb < 2; ++sb) {       // the residual of a buffer's next use is loaded as soon as the buffer is free
                    pre[sb] = cur;
                    while (pre[sb].tile < n_tiles && sb_of(pre[sb]) != sb) next(pre[sb]);
                    if (pre[sb].tile < n_tiles) load_res(pre[sb], sb);
                }
            uint32_t uses[2] = {0u, 0u};
            for (; cur.tile < n_tiles; next(cur)) {
                const uint32_t sb = sb_of(cur), sph = uses[sb]++ & 1u;
                int row0, col0, nblk;
                coords(cur, row0, col0, nblk);
                mbar_wait(gfull_bar(sb), sph);
                for (int j = 0; j < nblk; ++j) {
                    if (a.l2_flags & 2) tma_store_2d_hint(&ymap, col0 + 64 * j, row0, s_stg + sb * C::kUnitBytes + (uint32_t)j * (BLOCK_M * 128u), pol_last);
                    else tma_store_2d(&ymap, col0 + 64 * j, row0, s_stg + sb * C::kUnitBytes + (uint32_t)j * (BLOCK_M * 128u));
                }
                bulk_commit();
                bulk_wait_read0();          // the unit has been read out of shared memory
                if (!has_res) mbar_arrive(sfree_bar(sb));
                else {
                    next_same(pre[sb]);
                    if (pre[sb].tile < n_tiles) load_res(pre[sb], sb);
                }
            }
            bulk_wait_all();                // the last stores have reached global memory
        }
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;                       // warps 6..9 -> lane quarters 2,3,0,1
        const int row = q * 32 + lane;
        const int HoWo = a.Ho * a.Wo;
        TOut *y = static_cast<TOut *>(a.y);
        const TRes *res = static_cast<const TRes *>(a.res);
        const bool wide_st = sizeof(TOut) == 2 && (reinterpret_cast<uintptr_t>(a.y) & 31) == 0 && a.out_ld % 16 == 0 && a.out_choff % 16 == 0;
        const bool wide_f32 = sizeof(TOut) == 4 && (reinterpret_cast<uintptr_t>(a.y) & 31) == 0 && a.out_ld % 8 == 0 && a.out_choff % 8 == 0;
        (void)wide_f32;
        const int grp = warp >= 10 ? 1 : 0;
        uint32_t it = 0, uses0 = 0, uses1 = 0;       // uses of staging buffer 0 / 1 so far (N = 256: by either group)
        (void)uses0; (void)uses1;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            if (!kSplitCols && two_groups && (int)(it & 1u) != grp) continue;      // the groups take alternate tiles
            const long long mt = tile / n_tiles_n;
            const int nt = (int)(tile - mt * n_tiles_n);
            const long long m = mt * BLOCK_M + row;
            const int n0 = nt * BLOCK_N;
            const uint32_t buf = it & 1u, acc_phase = (it >> 1) & 1u;
            const float *grow = nullptr;
            if (a.gate != nullptr && m < a.M) {
                const int b = (int)(m / HoWo);
                const int wo = (int)(m % a.Wo);
                grow = a.gate + ((long long)b * a.gate_nwin + wo / a.gate_win) * a.Cout;
            }
            if constexpr (TEPI) {
                // ---- TMA epilogue: the tile leaves in units of <= 128 columns through two 128B-swizzled staging buffers
                // (one or two blocks of 128 rows x 64 columns each).  The store warp TMA-loads a residual unit into the
                // buffer beforehand (updated in place here) and TMA-stores the finished unit (rows past M / columns past
                // Cout are clipped), so this warp group never waits on a store: the per-tile chain is the epilogue math
                // alone.  (With one buffer and the store issued from here, every tile cost ~3.5 us however small:
                // math -> store -> wait for the read -> residual load -> next tile.)  Row-strided per-lane global accesses
                // - 32 different lines per warp instruction - are what made the residual GEMMs run below 1 TB/s before.
                if (warp == 6) GEMM_TS(it, 6);
                mbar_wait(accf_bar(buf), acc_phase);
                tc_fence_after();
                if (warp == 6) GEMM_TS(it, 7);
                const uint32_t taddr = tmem_base + buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
                const uint32_t x7 = (uint32_t)(row & 7);
                const int units = min(C::kUnitsPerTile, (a.Cout - n0 + 127) / 128);
                const int h_lo = kSplitCols ? grp : 0, h_hi = kSplitCols ? min(units, grp + 1) : units;
                for (int h = h_lo; h < h_hi; ++h) {
                const uint32_t sb = kSplitCols ? (uint32_t)h : (it & 1u);
                const uint32_t sph = kSplitCols ? ((h == 0 ? uses0 : uses1) & 1u) : ((it >> 1) & 1u);
                if (res != nullptr) mbar_wait(rfull_bar(sb), sph);
                else mbar_wait(sfree_bar(sb), sph ^ 1u);
                const uint32_t srow = s_stg + sb * C::kUnitBytes + (uint32_t)row * 128u;
                const int c_end = min(BLOCK_N, h * 128 + 128);
#pragma unroll 1
                for (int c0 = h * 128; c0 < c_end; c0 += 32) {
                    const int n = n0 + c0;
                    if (n >= a.Cout) break;
                    const bool second = n + 16 < a.Cout;
                    uint32_t r[32];
                    {
                        uint32_t (&r0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[0]);
                        uint32_t (&r1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[16]);
                        tmem_ld16(taddr + c0, r0);
                        tmem_ld16(taddr + c0 + 16, r1);
                    }
                    float4 s4[8], h4[8];
                    if (a.epi_scale != nullptr) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int ne = (e < 4 || second) ? n + 4 * e : n;
                            s4[e] = __ldg(reinterpret_cast<const float4 *>(a.epi_scale + ne));
                            h4[e] = __ldg(reinterpret_cast<const float4 *>(a.epi_shift + ne));
                        }
                    }
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
                    if (a.epi_scale != nullptr) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            v[4 * e] = fmaf(v[4 * e], s4[e].x, h4[e].x); v[4 * e + 1] = fmaf(v[4 * e + 1], s4[e].y, h4[e].y);
                            v[4 * e + 2] = fmaf(v[4 * e + 2], s4[e].z, h4[e].z); v[4 * e + 3] = fmaf(v[4 * e + 3], s4[e].w, h4[e].w);
                        }
                    }
                    // this step's four 16-byte chunks of the row inside block c0 / 64
                    const uint32_t blk = srow + (uint32_t)((c0 & 127) >> 6) * (BLOCK_M * 128u);
                    const uint32_t ch0 = (uint32_t)((c0 & 63) >> 3);
                    if (res != nullptr) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint4 t = lds16(blk + (((ch0 + e) ^ x7) << 4));
                            const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162 *>(&w4[h]);
                                v[8 * e + 2 * h] += __low2float(b2);
                                v[8 * e + 2 * h + 1] += __high2float(b2);
                            }
                        }
                    }
                    apply_act_vec(v, a.act);
                    if (a.post_scale != nullptr) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            if (e < 16 || second) {
                                const float4 p4 = __ldg(reinterpret_cast<const float4 *>(a.post_scale + n + e));
                                const float4 q4 = __ldg(reinterpret_cast<const float4 *>(a.post_shift + n + e));
                                v[e] = fmaf(v[e], p4.x, q4.x); v[e + 1] = fmaf(v[e + 1], p4.y, q4.y);
                                v[e + 2] = fmaf(v[e + 2], p4.z, q4.z); v[e + 3] = fmaf(v[e + 3], p4.w, q4.w);
                            }
                        }
                        apply_act_vec(v, a.post_act);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        sts16(blk + (((ch0 + e) ^ x7) << 4),
                              make_uint4(pack2(v[8 * e], v[8 * e + 1]), pack2(v[8 * e + 2], v[8 * e + 3]), pack2(v[8 * e + 4], v[8 * e + 5]),
                                         pack2(v[8 * e + 6], v[8 * e + 7])));
                }
                fence_proxy_async();            // staging writes (generic proxy) -> TMA store (async proxy)
                mbar_arrive(gfull_bar(sb));
                }
                ++uses0;
                if (units == 2) ++uses1;
                tc_fence_before();              // this group's TMEM reads of the tile are done
                mbar_arrive(acce_bar(buf));
                continue;
            }
            mbar_wait(accf_bar(buf), acc_phase);
            tc_fence_after();
            if (warp == 6) GEMM_TS(it, 6);
            const uint32_t taddr = tmem_base + buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
            // 32 accumulator columns per step: both TMEM loads and the scale/shift loads are in flight together
#pragma unroll 1
            for (int c0 = kSplitCols ? grp * 128 : 0; c0 < (kSplitCols ? grp * 128 + 128 : BLOCK_N); c0 += 32) {
                const int n = n0 + c0;
                if (n >= a.Cout) break;
                const bool second = n + 16 < a.Cout;         // Cout is a multiple of 16
                uint32_t r[32];
                {
                    uint32_t (&r0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[0]);
                    uint32_t (&r1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[16]);
                    tmem_ld16(taddr + c0, r0);
                    tmem_ld16(taddr + c0 + 16, r1);
                }
                float4 s4[8], h4[8];
                if (a.epi_scale != nullptr) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int ne = (e < 4 || second) ? n + 4 * e : n;
                        s4[e] = __ldg(reinterpret_cast<const float4 *>(a.epi_scale + ne));
                        h4[e] = __ldg(reinterpret_cast<const float4 *>(a.epi_shift + ne));
                    }
                }
                tmem_ld_wait();
                if (m < a.M) {
                    float v[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
                    if (a.epi_scale != nullptr) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            v[4 * e] = fmaf(v[4 * e], s4[e].x, h4[e].x); v[4 * e + 1] = fmaf(v[4 * e + 1], s4[e].y, h4[e].y);
                            v[4 * e + 2] = fmaf(v[4 * e + 2], s4[e].z, h4[e].z); v[4 * e + 3] = fmaf(v[4 * e + 3], s4[e].w, h4[e].w);
                        }
                    }
                    if (res != nullptr) {
                        const TRes *rp = res + m * a.res_ld + a.res_choff + n;
                        if constexpr (sizeof(TRes) == 2) {       // 16-byte loads (res_ld, res_choff are multiples of 8)
#pragma unroll
                            for (int e = 0; e < 32; e += 8) {
                                if (e < 16 || second) {
                                    const uint4 t = ldg16(rp + e);
                                    const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                                    for (int h = 0; h < 4; ++h) {
                                        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162 *>(&w4[h]);
                                        v[e + 2 * h] += __low2float(b2);
                                        v[e + 2 * h + 1] += __high2float(b2);
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; e += 4) {
                                if (e < 16 || second) {
                                    const float4 t = __ldg(reinterpret_cast<const float4 *>(rp + e));
                                    v[e] += t.x; v[e + 1] += t.y; v[e + 2] += t.z; v[e + 3] += t.w;
                                }
                            }
                        }
                    }
                    if (grow != nullptr && a.gate_additive) {       // per-segment bias before the activation
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            if (e < 16 || second) {
                                const float4 g4 = __ldg(reinterpret_cast<const float4 *>(grow + n + e));
                                v[e] += g4.x; v[e + 1] += g4.y; v[e + 2] += g4.z; v[e + 3] += g4.w;
                            }
                        }
                    }
                    apply_act_vec(v, a.act);
                    if (grow != nullptr && !a.gate_additive) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            if (e < 16 || second) {
                                const float4 g4 = __ldg(reinterpret_cast<const float4 *>(grow + n + e));
                                v[e] *= g4.x; v[e + 1] *= g4.y; v[e + 2] *= g4.z; v[e + 3] *= g4.w;
                            }
                        }
                    }
                    if (a.post_scale != nullptr) {                   // conv -> act -> BN (-> act)
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            if (e < 16 || second) {
                                const float4 p4 = __ldg(reinterpret_cast<const float4 *>(a.post_scale + n + e));
                                const float4 q4 = __ldg(reinterpret_cast<const float4 *>(a.post_shift + n + e));
                                v[e] = fmaf(v[e], p4.x, q4.x); v[e + 1] = fmaf(v[e + 1], p4.y, q4.y);
                                v[e + 2] = fmaf(v[e + 2], p4.z, q4.z); v[e + 3] = fmaf(v[e + 3], p4.w, q4.w);
                            }
                        }
                        apply_act_vec(v, a.post_act);
                    }
                    TOut *yp = y + m * a.out_ld + a.out_choff + n;
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        if (hb == 1 && !second) break;
                        const float *vv = v + 16 * hb;
                        TOut *yq = yp + 16 * hb;
                        if constexpr (sizeof(TOut) == 2) {
                            const uint4 lo = make_uint4(pack2(vv[0], vv[1]), pack2(vv[2], vv[3]), pack2(vv[4], vv[5]), pack2(vv[6], vv[7]));
                            const uint4 hi = make_uint4(pack2(vv[8], vv[9]), pack2(vv[10], vv[11]), pack2(vv[12], vv[13]), pack2(vv[14], vv[15]));
                            if (wide_st) {      // one full 32-byte sector per lane instead of two half-sector stores
                                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yq), "r"(lo.x), "r"(lo.y),
                                             "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
                            } else {
                                *reinterpret_cast<uint4 *>(yq) = lo;
                                *reinterpret_cast<uint4 *>(yq + 8) = hi;
                            }
                        } else if (wide_f32) {      // full 32-byte sectors per lane (two per 16 columns) instead of four 16-byte stores
#pragma unroll
                            for (int e = 0; e < 16; e += 8)
                                stg32(yq + e, make_uint4(__float_as_uint(vv[e]), __float_as_uint(vv[e + 1]), __float_as_uint(vv[e + 2]), __float_as_uint(vv[e + 3])),
                                      make_uint4(__float_as_uint(vv[e + 4]), __float_as_uint(vv[e + 5]), __float_as_uint(vv[e + 6]), __float_as_uint(vv[e + 7])));
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; e += 4)
                                *reinterpret_cast<float4 *>(yq + e) = make_float4(vv[e], vv[e + 1], vv[e + 2], vv[e + 3]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acce_bar(buf));
            if (warp == 6) GEMM_TS(it, 7);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// [rows, cols] bf16 matrix with row pitch ld (elements), K-major box {64, box_rows}, 128B swizzle
int make_map(const void *ptr, long long rows, int cols, long long ld, int box_rows, CUtensorMap *out) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SPK_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%d ld=%lld", (int)r, rows, cols, ld);
        return SPK_ERR_CUDA;
    }
    return SPK_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const int *, const int *, cuuint32_t, cuuint32_t, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeIm2colFn encode_im2col_fn() {
    static EncodeIm2colFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeIm2colFn>(p);
    }();
    return fn;
}

// IM2COL map over the channels-last activations [B][H][W][ld] (channel axis ending at the conv's last input channel):
// pixel box = filter base positions [-pad, size + pad - (k - 1) dil), traversal stride = conv stride, 64 channels x 128
// pixels per load, 128B swizzle
int make_im2col_map(const ConvArgs &a, CUtensorMap *out) {
    EncodeIm2colFn fn = encode_im2col_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeIm2col is not available from the driver");
        return SPK_ERR_CUDA;
    }
    const cuuint64_t dims[4] = {(cuuint64_t)(a.in_choff + a.Cin), (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    const cuuint64_t strides[3] = {(cuuint64_t)a.in_ld * 2, (cuuint64_t)a.W * a.in_ld * 2, (cuuint64_t)a.H * a.W * a.in_ld * 2};
    const int lower[2] = {-a.pw, -a.ph};
    const int upper[2] = {a.pw - (a.KW - 1) * a.dw, a.ph - (a.KH - 1) * a.dh};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)a.sw, (cuuint32_t)a.sh, 1u};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(a.x), dims, strides, lower, upper, (cuuint32_t)BLOCK_K,
                    (cuuint32_t)BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col failed (%d): [%d][%d][%d][%d] k%dx%d s%d,%d p%d,%d d%d,%d", (int)r, a.B, a.H, a.W, a.in_ld, a.KH,
                  a.KW, a.sh, a.sw, a.ph, a.pw, a.dh, a.dw);
        return SPK_ERR_CUDA;
    }
    // small-tensor workaround the CUTLASS host code applies for drivers up to 13.1 (cute/atom/copy_traits_sm90_im2col.hpp)
    int drv = 0;
    if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (long long)a.B * a.H * a.W * a.in_ld * 2 < 131072)
        reinterpret_cast<uint64_t *>(out)[1] &= ~(1ull << 21);
    return SPK_OK;
}

bool is_plain_gemm(const ConvArgs &a) {
    return a.KH == 1 && a.KW == 1 && a.sh == 1 && a.sw == 1 && a.ph == 0 && a.pw == 0 && a.Ho == a.H && a.Wo == a.W;
}

int bf16_vector(const float *src, int n, const bf16 **out, cudaStream_t s) {
    static std::mutex mu;
    static std::map<const float *, bf16 *> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(src);
    if (it != cache.end()) {
        *out = it->second;
        return SPK_OK;
    }
    bf16 *d = nullptr;
    SPK_CUDA_OK(cudaMalloc(&d, (size_t)(n + 8) * sizeof(bf16)));
    SPK_CUDA_OK(cudaMemsetAsync(d, 0, (size_t)(n + 8) * sizeof(bf16), s));
    int rc = launch_f32_to_bf16(src, d, n, s);
    if (rc != SPK_OK) return rc;
    cache[src] = d;
    *out = d;
    return SPK_OK;
}

template <int BLOCK_N, typename TOut, typename TRes, bool TEPI = false>
int launch_one(const ConvArgs &a, cudaStream_t s) {
    using C = Cfg<BLOCK_N, TEPI>;
    auto kern = conv_gemm_kernel<BLOCK_N, TOut, TRes, TEPI>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_gemm) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    CUtensorMap amap, wmap;
    const int im2col = is_plain_gemm(a) ? 0 : 1;
    int rc = im2col ? make_im2col_map(a, &amap) : make_map(static_cast<const bf16 *>(a.x) + a.in_choff, a.M, a.Cin, a.in_ld, BLOCK_M, &amap);
    if (rc == SPK_OK) rc = make_map(a.w, a.Cout, a.K, a.K, BLOCK_N, &wmap);
    if (rc != SPK_OK) return rc;
    CUtensorMap ymap = amap, rmap = amap;        // placeholders unless the TMA epilogue is used
    if (TEPI) {
        rc = make_map(static_cast<const bf16 *>(a.y) + a.out_choff, a.M, a.Cout, a.out_ld, BLOCK_M, &ymap);
        if (rc == SPK_OK && a.res != nullptr)
            rc = make_map(static_cast<const bf16 *>(a.res) + a.res_choff, a.M, a.Cout, a.res_ld, BLOCK_M, &rmap);
        if (rc != SPK_OK) return rc;
    }
    const bf16 *ps = nullptr, *ph = nullptr;
    if (a.pro_scale != nullptr) {
        if (a.pro_scale_bf != nullptr && a.pro_shift_bf != nullptr) {      // prepared by the model at set_program time
            ps = static_cast<const bf16 *>(a.pro_scale_bf);
            ph = static_cast<const bf16 *>(a.pro_shift_bf);
            rc = SPK_OK;
        } else {
            rc = bf16_vector(a.pro_scale, a.Cin, &ps, s);
            if (rc == SPK_OK) rc = bf16_vector(a.pro_shift, a.Cin, &ph, s);
        }
        if (rc != SPK_OK) return rc;
    }
    const long long mt = (a.M + BLOCK_M - 1) / BLOCK_M;
    const int ntn = (a.Cout + BLOCK_N - 1) / BLOCK_N;
    const long long tiles = mt * ntn;
    const long long grid = std::min<long long>(tiles, sm_count());
    static const int dbg_k = getenv("SPK_GEMM_DBG") ? atoi(getenv("SPK_GEMM_DBG")) : 0;      // timeline of launches with this K
    const int dbg = dbg_k != 0 && dbg_k == a.K;
    const int smem_bytes = C::kSmemBytes + (a.pro_scale != nullptr ? 2 * ((a.K + BLOCK_K - 1) / BLOCK_K) * 128 : 0);      // + prologue scale/shift
    if (smem_bytes > 227 * 1024) {
        set_error("conv_gemm: K=%d too large for the shared-memory prologue tables", a.K);
        return SPK_ERR_UNSUPPORTED;
    }
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), (size_t)smem_bytes, s, a, ps, ph, ntn, tiles, amap, wmap, ymap, rmap, dbg, im2col);
    if (le != cudaSuccess) {
        set_error("conv_gemm_kernel launch failed: %s", cudaGetErrorString(le));
        return SPK_ERR_CUDA;
    }
    return check_launch("conv_gemm_kernel");
}

template <typename TOut, typename TRes>
int launch_n(const ConvArgs &a, cudaStream_t s) {
    if constexpr (sizeof(TOut) == 2 && sizeof(TRes) == 2) {
        // bf16 tiles leave (and residual tiles arrive) through shared memory with TMA: faster than the register
        // epilogue for every GEMM of the three networks even with the shorter stage ring (CAM++ +4.6 %, ECAPA-TDNN
        // +12 %, ERes2NetV2 +23 % in its first, single-buffer form).  SPK_GEMM_TEPI=0 selects the register epilogue
        // for A/B runs.
        static const int force = [] { const char *e = getenv("SPK_GEMM_TEPI"); return e ? atoi(e) : -1; }();
        const bool ok = a.gate == nullptr && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0 &&
                        (a.res == nullptr || (reinterpret_cast<uintptr_t>(a.res) & 15) == 0);
        const bool want = force != 0;
        if (ok && want) {
            if (a.Cout <= 32) return launch_one<32, TOut, TRes, true>(a, s);
            if (a.Cout <= 64) return launch_one<64, TOut, TRes, true>(a, s);
            if (a.Cout <= 128) return launch_one<128, TOut, TRes, true>(a, s);
            return launch_one<256, TOut, TRes, true>(a, s);
        }
    }
    if (a.Cout <= 32) return launch_one<32, TOut, TRes>(a, s);
    if (a.Cout <= 64) return launch_one<64, TOut, TRes>(a, s);
    if (a.Cout <= 128) return launch_one<128, TOut, TRes>(a, s);
    return launch_one<256, TOut, TRes>(a, s);
}

}  // namespace

bool conv_gemm_supported(const ConvArgs &a, int in_dtype) {
    if (!conv_tc_supported(a, in_dtype)) return false;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.w) & 15) != 0) return false;
    if (is_plain_gemm(a)) return true;
    // everything else: TMA im2col loads (zero padding only, no per-channel prologue in front of the padding)
    static const bool off = [] { const char *e = getenv("SPK_NO_IM2COL"); return e && e[0] == '1'; }();
    if (off || a.pro_scale != nullptr) return false;
    if (a.pad_reflect && !reflect_edge_fix_supported(a)) return false;      // zero-padded conv, then the edge fix
    if (a.Cin % 8 != 0 || a.sh > 8 || a.sw > 8 || a.in_ld % 8 != 0 || a.in_choff % 8 != 0) return false;
    const int lw = -a.pw, lh = -a.ph, uw = a.pw - (a.KW - 1) * a.dw, uh = a.ph - (a.KH - 1) * a.dh;
    if (lw < -128 || lh < -128 || uw < -128 || uh < -128 || uw > 127 || uh > 127) return false;
    if ((a.KW - 1) * a.dw > 255 || (a.KH - 1) * a.dh > 255) return false;
    if (a.W + uw - lw < 1 || a.H + uh - lh < 1) return false;            // the pixel box must not be empty
    if (a.Wo != (a.W + uw - lw - 1) / a.sw + 1 || a.Ho != (a.H + uh - lh - 1) / a.sh + 1) return false;
    return true;
}

int launch_conv_gemm_main(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s);

int launch_conv_gemm(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s) {
    const int rc = launch_conv_gemm_main(a, out_dtype, res_dtype, s);
    if (rc != SPK_OK || !a.pad_reflect || a.M == 0) return rc;
    return launch_reflect_edge_fix(a, SPK_DT_BF16, out_dtype, s);
}

int launch_conv_gemm_main(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s) {
    if (a.M == 0) return SPK_OK;
    const bool res_bf16 = a.res == nullptr ? (out_dtype == SPK_DT_BF16) : (res_dtype == SPK_DT_BF16);
    if (out_dtype == SPK_DT_BF16) {
        if (!res_bf16) {
            set_error("conv_gemm: a bf16 output takes a bf16 residual");
            return SPK_ERR_UNSUPPORTED;
        }
        return launch_n<bf16, bf16>(a, s);
    }
    return res_bf16 ? launch_n<float, bf16>(a, s) : launch_n<float, float>(a, s);
}

}  // namespace spk

// debug aid (not part of the ABI)
extern "C" int spk_debug_gemm_timeline(long long *dst) {
    return cudaMemcpyFromSymbol(dst, spk::g_gemm_ts, sizeof(long long) * 256 * 8) == cudaSuccess ? 0 : -1;
}
