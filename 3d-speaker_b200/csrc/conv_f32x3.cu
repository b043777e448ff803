// fp32-accurate convolution GEMM on the tensor cores (3xTF32): the "fp32" precision mode of the embedding networks
// without leaving tcgen05.  That mode used to run every conv on CUDA cores (conv_simt.cu); here the fp32 operands are
// split into a TF32 head and a TF32 tail,
//
//   x = x_hi + x_lo,  x_hi = x rounded to TF32,  x_lo = x - x_hi  (exact in fp32)
//   D = A_lo * W_hi + A_hi * W_lo + A_hi * W_hi              (the A_lo * W_lo term, 2^-22 relative, is dropped)
//
// and the three products go through kind::tf32 MMAs.  The TMEM accumulator ADDS WITH TRUNCATION (measured: a K-long sum
// left in TMEM comes out -5.9e-6 relative at K = 992, -2.5e-5 at K = 4096, always towards zero, coherent from layer to
// layer), so TMEM only ever holds the partial sum of one 32-deep chunk: drain warps add every chunk into fp32 REGISTER
// sums with round-to-nearest (Ootomo & Yokota's remedy) while the next chunks fill the other buffers of a TMEM ring.
// Result: 1.4e-7 .. 2.5e-7 rel-L2 against float64 at any K - tighter than the CUDA-core fp32 sum (2e-7 .. 6e-7) - and
// embeddings within 1.1e-6 rel-L2 of the CUDA-core path where the literal tolerance against the reference is 1e-4.
//
// Data path: activations stay fp32 [M, ld] channels-last in HBM.  TMA lands a 128 x 32 fp32 tile of A (plain 2-D tiles
// for stride-1 1x1 convs, TMA IM2COL loads of 32 channels x 128 pixels per filter tap otherwise, as in conv_gemm.cu)
// and the matching tiles of W_hi and W_lo (split once per model at spk_model_set_program).  Split warps (thread = tile
// row = TMEM lane) apply the optional BN-ReLU prologue in fp32, split the landed row and store head and tail to
// TENSOR MEMORY (tcgen05.st); the MMAs take A from there.  (With head and tail written back to shared memory and
// re-read by every MMA a stage moved 132 KB (N = 32) to 192 KB (N = 128) through shared memory: 1000-1500 cycles at
// 128 B/clk against 200-770 cycles of tensor-core work.)  Epilogue from the register sums: folded BN, residual, CAM
// gate, activation, ECAPA's post-affine.
//
// What bounds it (role timeline, SPK_F32X3_DBG + tools/x3_timeline.py): the single-thread roles.  One thread issuing
// the 12 MMAs, three commits and two barrier waits of a stage runs a ~150-instruction dependent chain, ~1000 cycles
// per stage; one thread issuing the three tensor loads of a stage ~780 cycles.  Hence two MMA-issuing warps (chunks
// are independent accumulators: alternate chunks), two producer warps (activation / weight tiles), two split groups
// on alternate stages for N <= 64, and a division-free split loop with integer TF32 rounding (cvt.rna.tf32 compiles
// to four instructions).  First FCM conv of CAM++ (32 -> 32 channels, 3x3, 256 segments): 448 -> 313 us.
//
// Reference ops served: every nn.Conv1d / nn.Conv2d of the fp32 eval forward of CAM++ (speakerlab/models/campplus/
// layers.py, DTDNN.py), ERes2Net / ERes2NetV2 (speakerlab/models/eres2net/) and ECAPA-TDNN (speakerlab/models/
// ecapa_tdnn/ECAPA_TDNN.py); reflect-padded Conv1d's run zero-padded here and have their edge positions recomputed by
// reflect_edge_fix_kernel (small_ops.cu); the per-segment dense layers stay on the split-K linear kernel.
#include <cstdlib>
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "ops.cuh"
#include "tc.cuh"
#include "tmap.cuh"

namespace spk {
namespace {

using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;        // fp32 elements per stage = one 128-byte swizzled row
constexpr int UMMA_K = 8;          // kind::tf32
// warps: 0 and 15 TMA (activation / weight tiles), 1 and 14 MMA issue (alternate chunks), 2-5 split/prologue, 6-9 drain/epilogue, 10-13 second drain group (N = 128: one column half each) or
// second split group (N <= 64: alternate stages)
constexpr int kXformThreads = 128;
constexpr int kEpilogueThreads = 128;

// A (the split tile) from tensor memory: lane = row, one 32-bit column per K element
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap *map, int c, int w, int h, int n, uint16_t off_w,
                                                   uint16_t off_h, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ uint32_t sw128_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kSw128Hi = (1024u >> 4) | (1u << 14) | (2u << 29);
// fp32 accumulate, TF32 x TF32 (format code 2), K-major both, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}
__device__ __forceinline__ float tf32_head(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// debug aid: per-stage role timestamps of CTA 0 (SPK_F32X3_DBG=1), read back by spk_debug_f32x3_timeline
__device__ long long g_x3_ts[256 * 8];
#define X3_TS(idx, slot) do { if (dbg && blockIdx.x == 0 && (idx) >= dbg_from && (idx) < dbg_from + 256 && (threadIdx.x & 31) == 0) g_x3_ts[((idx) - dbg_from) * 8 + (slot)] = clock64(); } while (0)

template <int BLOCK_N> struct Cfg3 {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 4;       // the landed fp32 A tile
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 4;       // one of the two W tiles (heads, tails)
    static constexpr int kStageBytes = kABytes + 2 * kBBytes;
    static constexpr int kStagesFit = ((227 * 1024 - 2048) / kStageBytes) & ~1;      // even: see kXformGroups
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kXformGroups = BLOCK_N <= 64 ? 2 : 1;  // split groups on alternate stages (short MMA phases need both)
    static constexpr int kDrainGroups = BLOCK_N <= 64 ? 1 : 2;  // drain threads keep <= 64 running sums each
    static constexpr int kThreads = 512;                        // warp 14: second MMA issuer, warp 15: weight-tile producer
    static constexpr int kSlots = 4;                            // TMEM ring of split A tiles: 32 head + 32 tail columns each
    static constexpr int kSlotCols = 2 * BLOCK_K;
    static constexpr int kAccBufs = BLOCK_N <= 64 ? 4 : 2;      // TMEM accumulator ring: chunk partial sums waiting for their drain
    static constexpr int kAccCols = kAccBufs * BLOCK_N;
    static constexpr int kTmemNeed = kAccCols + kSlots * kSlotCols;
    static constexpr int kTmemCols = kTmemNeed <= 256 ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512;
    static_assert(kStages >= 2 && kStages % 2 == 0 && kTmemNeed <= 512, "ring sizes");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(Cfg3<BLOCK_N>::kThreads, 1)
conv_f32x3_kernel(const ConvArgs a, int n_tiles_n, long long n_tiles, const __grid_constant__ CUtensorMap amap,
                  const __grid_constant__ CUtensorMap whmap, const __grid_constant__ CUtensorMap wlmap, int im2col, int chunk_stages, int dbg) {
    const long long dbg_from = 512;          // skip the ramp
    using C = Cfg3<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (base - raw) + C::kStages * C::kStageBytes);
    const uint32_t bar0 = smem_u32(bars);
    auto land_bar = [&](int s) { return bar0 + 8u * s; };
    auto ready_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + s); };
    auto accf_bar = [&](int b) { return bar0 + 8u * (3 * C::kStages + b); };
    auto acce_bar = [&](int b) { return bar0 + 8u * (3 * C::kStages + C::kAccBufs + b); };
    auto tfree_bar = [&](int t) { return bar0 + 8u * (3 * C::kStages + 2 * C::kAccBufs + t); };      // split slot t read by its MMAs
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bars + 3 * C::kStages + 2 * C::kAccBufs + C::kSlots);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(land_bar(s), 1);
            mbar_init(ready_bar(s), kXformThreads);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < C::kAccBufs; ++b) {
            mbar_init(accf_bar(b), 1);
            mbar_init(acce_bar(b), C::kDrainGroups * kEpilogueThreads);      // every drain group reads every chunk
        }
        for (int t = 0; t < C::kSlots; ++t) mbar_init(tfree_bar(t), 1);
        fence_barrier_init();
        prefetch_tmap(&amap);
        prefetch_tmap(&whmap);
        prefetch_tmap(&wlmap);
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int chunks = (a.Cin + BLOCK_K - 1) / BLOCK_K;                      // im2col: 32-channel chunks per filter tap
    const int nk = im2col ? a.KH * a.KW * chunks : (a.K + BLOCK_K - 1) / BLOCK_K;
    pdl_wait();

    if (warp == 0 || warp == 15) {
        // =========================== TMA producers ===========================
        // Issuing a tensor load costs the issuing thread a few hundred cycles; with the activation tile and both weight
        // tiles issued by one thread the producer loop (~780 cycles per stage) paced the whole kernel.  Warp 0 issues the
        // activation tiles (and arms the stage barrier with the stage's full byte count), warp 15 the weight tiles.
        if (lane == 0) {
            const bool act = warp == 0;
            int stage = 0;
            long long sidx = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const long long mt = tile / n_tiles_n;
                const int nt = (int)(tile - mt * n_tiles_n);
                int bw = 0, bh = 0, bn = 0, kh = 0, kw = 0, ch = 0;
                if (im2col && act) {          // base position (filter tap 0, 0) of the tile's first output pixel
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
                    X3_TS(sidx, act ? 1 : 2);
                    const uint32_t sa = base + stage * C::kStageBytes;
                    if (act) {
                        mbar_arrive_expect_tx(land_bar(stage), C::kStageBytes);
                        if (im2col) {
                            tma_load_im2col_4d(sa, &amap, a.in_choff + ch * BLOCK_K, bw, bh, bn, (uint16_t)(kw * a.dw), (uint16_t)(kh * a.dh),
                                               land_bar(stage));
                            if (++ch == chunks) { ch = 0; if (++kw == a.KW) { kw = 0; ++kh; } }
                        } else {
                            tma_load_2d(sa, &amap, kc * BLOCK_K, (int)(mt * BLOCK_M), land_bar(stage));
                        }
                    } else {
                        // weight column of this stage: tap-major packed K ((kh * KW + kw) * Cin + channel chunk)
                        int kcol = kc * BLOCK_K;
                        if (im2col) {
                            kcol = (kh * a.KW + kw) * a.Cin + ch * BLOCK_K;
                            if (++ch == chunks) { ch = 0; if (++kw == a.KW) { kw = 0; ++kh; } }
                        }
                        const uint32_t sbh = sa + C::kABytes, sbl = sbh + C::kBBytes;
                        tma_load_2d(sbh, &whmap, kcol, nt * BLOCK_N, land_bar(stage));
                        tma_load_2d(sbl, &wlmap, kcol, nt * BLOCK_N, land_bar(stage));
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 || warp == 14) {
        // =========================== MMA issuers ===========================
        // One issuing thread runs a serial chain of ~150 dependent instructions per stage (two barrier waits, fences, twelve
        // descriptor builds + MMAs, three commits): ~1000 cycles measured, five times the tensor-core time of an N = 32
        // stage.  Chunks accumulate in different TMEM buffers and are independent, so two warps issue alternate chunks.
        constexpr uint32_t idesc = make_idesc_tf32(BLOCK_N);
        const uint32_t mine_parity = warp == 1 ? 0u : 1u;
        int stage = 0, slot = 0, in_chunk = 0;
        long long sidx = 0;
        uint32_t phase = 0, gc = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < nk; ++kc, ++sidx) {
                const bool last = in_chunk == chunk_stages - 1 || kc == nk - 1;
                if ((gc & 1u) == mine_parity) {
                    const uint32_t buf = gc % C::kAccBufs, acc_phase = (gc / C::kAccBufs) & 1u;
                    X3_TS(sidx, 0);
                    if (in_chunk == 0) {               // a new chunk starts on the next TMEM buffer once that has been drained
                        mbar_wait(acce_bar(buf), acc_phase ^ 1u);
                        X3_TS(sidx, 7);
                        tc_fence_after();
                    }
                    const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
                    X3_TS(sidx, 4);
                    mbar_wait(ready_bar(stage), phase);          // split tile in TMEM (and with it the stage's W tiles landed)
                    tc_fence_after();
                    X3_TS(sidx, 5);
                    if (elect_one()) {
                        const uint32_t sa = base + stage * C::kStageBytes;
                        const uint32_t bh = sw128_lo(sa + C::kABytes), bl = sw128_lo(sa + C::kABytes + C::kBBytes);
                        const uint32_t th = tmem_base + C::kAccCols + slot * C::kSlotCols, tl = th + BLOCK_K;
                        // small terms first, then the head product
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                            umma_tf32_ta(d_tmem, tl + kk * UMMA_K, desc64(bh + 2u * kk, kSw128Hi), idesc, (in_chunk | kk) != 0 ? 1u : 0u);
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                            umma_tf32_ta(d_tmem, th + kk * UMMA_K, desc64(bl + 2u * kk, kSw128Hi), idesc, 1u);
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                            umma_tf32_ta(d_tmem, th + kk * UMMA_K, desc64(bh + 2u * kk, kSw128Hi), idesc, 1u);
                        umma_commit(empty_bar(stage));
                        umma_commit(tfree_bar(slot));
                        if (last) umma_commit(accf_bar(buf));
                    }
                    __syncwarp();
                    X3_TS(sidx, 6);
                }
                if (last) { ++gc; in_chunk = 0; } else ++in_chunk;
                if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                if (++slot == C::kSlots) slot = 0;
            }
        }
    } else if (warp < 6 || (warp >= 10 && C::kXformGroups == 2)) {
        // =========================== split (+ BN-ReLU prologue) into tensor memory ===========================
        // thread = tile row = TMEM lane: read the row's eight 16-byte chunks (128B-swizzled), apply the prologue, split,
        // store 32 heads and 32 tails as columns of this lane in the stage's TMEM slot.  With two groups (N <= 64) the
        // groups take alternate stage uses; rings are even, so a stage / slot always belongs to the same group.
        const int q = warp & 3, r = q * 32 + lane;
        const int grp = warp >= 10 ? 1 : 0;
        const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128), x7 = (uint32_t)(r & 7);
        const bool has_pro = a.pro_scale != nullptr, relu = a.pro_relu != 0;
        const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const long long n_uses = my_tiles * nk;
        uint32_t coff[8];                              // the row's eight 16-byte chunks, swizzled
#pragma unroll
        for (int j = 0; j < 8; ++j) coff[j] = row_off + (((uint32_t)j ^ x7) << 4);
        // ring positions advance by the group count (no 64-bit divisions in the loop: this role is instruction-bound)
        int stage = grp, slot = grp, kc = grp % nk;
        uint32_t phase = 0, tphase = 0;
        for (long long sidx = grp; sidx < n_uses; sidx += C::kXformGroups) {
            mbar_wait(tfree_bar(slot), tphase ^ 1u);          // the MMAs of the slot's previous tile are done
            tc_fence_after();
            mbar_wait(land_bar(stage), phase);
            const uint32_t sa = base + stage * C::kStageBytes;
            const uint32_t taddr = tmem_base + C::kAccCols + slot * C::kSlotCols + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = half * 4 + jj;
                    const uint4 v = lds16(sa + coff[j]);
                    float x[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
                    if (has_pro) {                 // plain GEMM only: K = Cin, a multiple of 4; warp-uniform (broadcast) loads
                        const int c = kc * BLOCK_K + j * 4;
                        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), h4 = s4;
                        if (c < a.K) {
                            s4 = __ldg(reinterpret_cast<const float4 *>(a.pro_scale + c));
                            h4 = __ldg(reinterpret_cast<const float4 *>(a.pro_shift + c));
                        }
                        x[0] = fmaf(x[0], s4.x, h4.x); x[1] = fmaf(x[1], s4.y, h4.y);
                        x[2] = fmaf(x[2], s4.z, h4.z); x[3] = fmaf(x[3], s4.w, h4.w);
                        if (relu) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
                        }
                    }
                    // head = the value rounded to TF32 with integer ops (add half an ulp, mask: two instructions where
                    // cvt.rna.tf32 compiles to four with its Inf/NaN guard - this role is instruction-bound; plain truncation,
                    // one instruction, leaves a -2.3e-7 bias because the MMA then truncates a same-signed 13-bit tail)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t h = (__float_as_uint(x[e]) + 0x1000u) & 0xFFFFE000u;
                        hi[4 * jj + e] = h;
                        lo[4 * jj + e] = __float_as_uint(x[e] - __uint_as_float(h));
                    }
                }
                tmem_st16(taddr + 16 * half, hi);
                tmem_st16(taddr + BLOCK_K + 16 * half, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(ready_bar(stage));
            if (q == 2) X3_TS(sidx, 3);
            stage += C::kXformGroups;
            if (stage >= C::kStages) { stage -= C::kStages; phase ^= 1u; }
            slot += C::kXformGroups;
            if (slot >= C::kSlots) { slot -= C::kSlots; tphase ^= 1u; }
            kc += C::kXformGroups;
            while (kc >= nk) kc -= nk;
        }
    } else {
        // =========================== drain + epilogue ===========================
        // The tensor core's accumulator adds with truncation: a K-long sum kept in TMEM shrinks towards zero by ~K/8 x 3
        // x 2^-26 relative (measured: -5.9e-6 at K = 992), coherently from layer to layer.  So TMEM only ever holds the
        // partial sum of one chunk (`chunk_stages` stages of 32; small products issued first); these warps drain every
        // chunk into fp32 REGISTER sums with round-to-nearest adds while the MMAs of the next chunk fill the other
        // TMEM buffer (Ootomo & Yokota's remedy for tensor-core accumulation), and run the conv epilogue from registers.
        constexpr int kCols = BLOCK_N / C::kDrainGroups;      // columns per thread (group g: [g * kCols, (g + 1) * kCols))
        constexpr int kPieces = kCols / 16;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int HoWo = a.Ho * a.Wo;
        float *y = static_cast<float *>(a.y);
        const float *res = static_cast<const float *>(a.res);
        const bool wide_f32 = (reinterpret_cast<uintptr_t>(a.y) & 31) == 0 && a.out_ld % 8 == 0 && a.out_choff % 8 == 0;
        const int grp = warp >= 10 ? 1 : 0;
        const int n_chunks = (nk + chunk_stages - 1) / chunk_stages;
        uint32_t gc = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long mt = tile / n_tiles_n;
            const int nt = (int)(tile - mt * n_tiles_n);
            const long long m = mt * BLOCK_M + row;
            const int n0 = nt * BLOCK_N + grp * kCols;
            float acc[kCols];
            for (int c = 0; c < n_chunks; ++c, ++gc) {
                const uint32_t buf = gc % C::kAccBufs, acc_phase = (gc / C::kAccBufs) & 1u;
                mbar_wait(accf_bar(buf), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + buf * BLOCK_N + grp * kCols + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int p = 0; p < kPieces; ++p) {             // the accumulator ring hides this round trip
                    uint32_t r0[16];
                    tmem_ld16(taddr + 16 * p, r0);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[16 * p + e] = c == 0 ? __uint_as_float(r0[e]) : acc[16 * p + e] + __uint_as_float(r0[e]);
                }
                tc_fence_before();
                mbar_arrive(acce_bar(buf));

            }
            if (m >= a.M) continue;
            const float *grow = nullptr;
            if (a.gate != nullptr) {
                const int b = (int)(m / HoWo);
                const int wo = (int)(m % a.Wo);
                grow = a.gate + ((long long)b * a.gate_nwin + wo / a.gate_win) * a.Cout;
            }
#pragma unroll
            for (int p = 0; p < kPieces; ++p) {
                const int n = n0 + 16 * p;
                if (n >= a.Cout) break;                  // Cout is a multiple of 16
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = acc[16 * p + e];
                if (a.epi_scale != nullptr) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 s4 = __ldg(reinterpret_cast<const float4 *>(a.epi_scale + n + e));
                        const float4 h4 = __ldg(reinterpret_cast<const float4 *>(a.epi_shift + n + e));
                        v[e] = fmaf(v[e], s4.x, h4.x); v[e + 1] = fmaf(v[e + 1], s4.y, h4.y);
                        v[e + 2] = fmaf(v[e + 2], s4.z, h4.z); v[e + 3] = fmaf(v[e + 3], s4.w, h4.w);
                    }
                }
                if (res != nullptr) {
                    const float *rp = res + m * a.res_ld + a.res_choff + n;
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 t4 = __ldg(reinterpret_cast<const float4 *>(rp + e));
                        v[e] += t4.x; v[e + 1] += t4.y; v[e + 2] += t4.z; v[e + 3] += t4.w;
                    }
                }
                if (grow != nullptr && a.gate_additive) {       // per-segment bias before the activation
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(grow + n + e));
                        v[e] += g4.x; v[e + 1] += g4.y; v[e + 2] += g4.z; v[e + 3] += g4.w;
                    }
                }
                apply_act_vec(v, a.act);
                if (grow != nullptr && !a.gate_additive) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(grow + n + e));
                        v[e] *= g4.x; v[e + 1] *= g4.y; v[e + 2] *= g4.z; v[e + 3] *= g4.w;
                    }
                }
                if (a.post_scale != nullptr) {                   // conv -> act -> BN (-> act)
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 p4 = __ldg(reinterpret_cast<const float4 *>(a.post_scale + n + e));
                        const float4 q4 = __ldg(reinterpret_cast<const float4 *>(a.post_shift + n + e));
                        v[e] = fmaf(v[e], p4.x, q4.x); v[e + 1] = fmaf(v[e + 1], p4.y, q4.y);
                        v[e + 2] = fmaf(v[e + 2], p4.z, q4.z); v[e + 3] = fmaf(v[e + 3], p4.w, q4.w);
                    }
                    apply_act_vec(v, a.post_act);
                }
                float *yq = y + m * a.out_ld + a.out_choff + n;
                if (wide_f32) {          // full 32-byte sectors per lane
#pragma unroll
                    for (int e = 0; e < 16; e += 8)
                        stg32(yq + e, make_uint4(__float_as_uint(v[e]), __float_as_uint(v[e + 1]), __float_as_uint(v[e + 2]), __float_as_uint(v[e + 3])),
                              make_uint4(__float_as_uint(v[e + 4]), __float_as_uint(v[e + 5]), __float_as_uint(v[e + 6]), __float_as_uint(v[e + 7])));
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        *reinterpret_cast<float4 *>(yq + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

__global__ void split_tf32_kernel(const float *__restrict__ src, float *__restrict__ hi, float *__restrict__ lo, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = src[i], h = tf32_head(x);
        hi[i] = h;
        lo[i] = x - h;
    }
}

// ------------------------------------------------------------------ host
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

// [rows, cols] fp32 matrix with row pitch ld (elements), K-major box {32, box_rows}, 128B swizzle
int make_map_f32(const void *ptr, long long rows, int cols, long long ld, int box_rows, CUtensorMap *out) {
    TmapEncodeTiledFn fn = tmap_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SPK_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (f32) failed (%d): rows=%lld cols=%d ld=%lld", (int)r, rows, cols, ld);
        return SPK_ERR_CUDA;
    }
    return SPK_OK;
}

// IM2COL map over fp32 channels-last activations [B][H][W][ld]: 32 channels x 128 pixels per load (see conv_gemm.cu)
int make_im2col_map_f32(const ConvArgs &a, CUtensorMap *out) {
    EncodeIm2colFn fn = encode_im2col_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeIm2col is not available from the driver");
        return SPK_ERR_CUDA;
    }
    const cuuint64_t dims[4] = {(cuuint64_t)(a.in_choff + a.Cin), (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    const cuuint64_t strides[3] = {(cuuint64_t)a.in_ld * 4, (cuuint64_t)a.W * a.in_ld * 4, (cuuint64_t)a.H * a.W * a.in_ld * 4};
    const int lower[2] = {-a.pw, -a.ph};
    const int upper[2] = {a.pw - (a.KW - 1) * a.dw, a.ph - (a.KH - 1) * a.dh};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)a.sw, (cuuint32_t)a.sh, 1u};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(a.x), dims, strides, lower, upper, (cuuint32_t)BLOCK_K,
                    (cuuint32_t)BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col (f32) failed (%d): [%d][%d][%d][%d] k%dx%d s%d,%d p%d,%d d%d,%d", (int)r, a.B, a.H, a.W, a.in_ld,
                  a.KH, a.KW, a.sh, a.sw, a.ph, a.pw, a.dh, a.dw);
        return SPK_ERR_CUDA;
    }
    int drv = 0;      // small-tensor workaround, as in conv_gemm.cu
    if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (long long)a.B * a.H * a.W * a.in_ld * 4 < 131072)
        reinterpret_cast<uint64_t *>(out)[1] &= ~(1ull << 21);
    return SPK_OK;
}

bool is_plain(const ConvArgs &a) {
    return a.KH == 1 && a.KW == 1 && a.sh == 1 && a.sw == 1 && a.ph == 0 && a.pw == 0 && a.Ho == a.H && a.Wo == a.W;
}

template <int BLOCK_N>
int launch_one(const ConvArgs &a, cudaStream_t s) {
    using C = Cfg3<BLOCK_N>;
    auto kern = conv_f32x3_kernel<BLOCK_N>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_f32x3) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    CUtensorMap amap, whmap, wlmap;
    const int im2col = is_plain(a) ? 0 : 1;
    const float *w_hi = static_cast<const float *>(a.w), *w_lo = w_hi + (long long)a.Cout * a.K;
    int rc = im2col ? make_im2col_map_f32(a, &amap) : make_map_f32(static_cast<const float *>(a.x) + a.in_choff, a.M, a.Cin, a.in_ld, BLOCK_M, &amap);
    if (rc == SPK_OK) rc = make_map_f32(w_hi, a.Cout, a.K, a.K, BLOCK_N, &whmap);
    if (rc == SPK_OK) rc = make_map_f32(w_lo, a.Cout, a.K, a.K, BLOCK_N, &wlmap);
    if (rc != SPK_OK) return rc;
    const long long mt = (a.M + BLOCK_M - 1) / BLOCK_M;
    const int ntn = (a.Cout + BLOCK_N - 1) / BLOCK_N;
    const long long tiles = mt * ntn;
    const long long grid = std::min<long long>(tiles, sm_count());
    // stages (of 32 K elements) summed inside TMEM before a drain: 1 = fp32-class sums, larger = fewer drains
    static const int chunk_stages = [] { const char *e = getenv("SPK_F32X3_CHUNK"); return e && atoi(e) > 0 ? atoi(e) : 1; }();
    static const int dbg_k = [] { const char *e = getenv("SPK_F32X3_DBG"); return e ? atoi(e) : 0; }();      // timeline of launches with this K
    const int dbg = dbg_k != 0 && dbg_k == a.K;
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3(C::kThreads), (size_t)C::kSmemBytes, s, a, ntn, tiles, amap, whmap, wlmap, im2col, chunk_stages, dbg);
    if (le != cudaSuccess) {
        set_error("conv_f32x3_kernel launch failed: %s", cudaGetErrorString(le));
        return SPK_ERR_CUDA;
    }
    return check_launch("conv_f32x3_kernel");
}

}  // namespace

bool conv_f32x3_supported(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype) {
    static const bool off = [] { const char *e = getenv("SPK_NO_F32X3"); return e && e[0] == '1'; }();
    if (off) return false;
    if (in_dtype != SPK_DT_F32 || out_dtype != SPK_DT_F32 || (a.res != nullptr && res_dtype != SPK_DT_F32)) return false;
    // per-segment dense layers (one output pixel per segment) stay on CUDA cores.  The rule must not
    // depend on the batch size: a segment's embedding may not change with the sub-batch it travels in
    if (a.Ho * a.Wo < 8) return false;
    if (a.pad_reflect && !reflect_edge_fix_supported(a)) return false;      // zero-padded conv here, then the edge fix (model.cu)
    if (a.Cin % 4 != 0 || a.in_ld % 4 != 0 || a.in_choff % 4 != 0 || a.K % 4 != 0) return false;
    if (a.Cout % 16 != 0 || a.out_ld % 4 != 0 || a.out_choff % 4 != 0) return false;
    if (a.res != nullptr && (a.res_ld % 4 != 0 || a.res_choff % 4 != 0 || (reinterpret_cast<uintptr_t>(a.res) & 15) != 0)) return false;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.y) & 15) != 0) return false;
    if (is_plain(a)) return true;
    if (a.pro_scale != nullptr) return false;                  // zero padding comes before the BN in the reference
    if (a.sh > 8 || a.sw > 8) return false;
    const int lw = -a.pw, lh = -a.ph, uw = a.pw - (a.KW - 1) * a.dw, uh = a.ph - (a.KH - 1) * a.dh;
    if (lw < -128 || lh < -128 || uw < -128 || uh < -128 || uw > 127 || uh > 127) return false;
    if ((a.KW - 1) * a.dw > 255 || (a.KH - 1) * a.dh > 255) return false;
    if (a.W + uw - lw < 1 || a.H + uh - lh < 1) return false;
    if (a.Wo != (a.W + uw - lw - 1) / a.sw + 1 || a.Ho != (a.H + uh - lh - 1) / a.sh + 1) return false;
    return true;
}

// a.w: [2][Cout][K] fp32 = TF32 heads then tails of the packed weights (launch_split_tf32)
int launch_conv_f32x3(const ConvArgs &a, cudaStream_t s) {
    if (a.M == 0) return SPK_OK;
    if (a.Cout <= 32) return launch_one<32>(a, s);
    if (a.Cout <= 64) return launch_one<64>(a, s);
    return launch_one<128>(a, s);          // wider outputs: 128-column tiles (64 running sums per drain thread)
}

int launch_split_tf32(const float *src, float *hi, float *lo, long long n, cudaStream_t s) {
    if (n <= 0) return SPK_OK;
    const int grid = (int)std::min<long long>((n + 255) / 256, 4096);
    split_tf32_kernel<<<grid, 256, 0, s>>>(src, hi, lo, n);
    return check_launch("split_tf32_kernel");
}

}  // namespace spk

// debug aid (not part of the ABI)
extern "C" int spk_debug_f32x3_timeline(long long *dst) {
    return cudaMemcpyFromSymbol(dst, spk::g_x3_ts, sizeof(long long) * 256 * 8) == cudaSuccess ? 0 : -1;
}
