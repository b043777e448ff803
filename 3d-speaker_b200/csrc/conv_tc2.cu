// Generic tcgen05 / TMEM implicit-GEMM convolution: the bf16 path of every conv shape that has no dedicated
// kernel (strided / dilated / multi-row kernels such as the CAM++ tdnn layer and the ERes2NetV2 3x3 convs).
//
// D[m,n] = sum_k A[m,k] W[n,k], A gathered on the fly from the channels-last input (zero or reflect padding)
// with optional BN-ReLU prologue; folded-BN / residual / activation / gate / post-activation-affine epilogue.
// The producer side is what a first ncu pass of a simpler gather kernel showed to matter (one producer warp per
// scheduler was instruction-latency bound: ~700 dependent instructions per 64-wide K chunk, register spills):
//   * W tiles come in by TMA (cp.async.bulk.tensor.2d, 128B swizzle, OOB rows/cols zero-filled)
//     issued by one producer lane with mbarrier expect_tx - no registers, no LSU instructions;
//   * 8 producer warps (two per scheduler) gather A; every thread owns 4 rows of one 16-byte
//     column and keeps the NEXT chunk's four 16-byte loads in flight while it finishes the
//     current one (register double buffering, across tile boundaries too);
//   * the BN-ReLU prologue runs in packed bf16x2 arithmetic (HFMA2.BF16 + HMNMX2: 8 instructions
//     per 16-byte piece instead of ~32), scale/shift pre-rounded to bf16;
//   * one persistent CTA per SM (416 threads: 8 producer, 4 epilogue, 1 MMA warp) with a 4-6
//     deep stage ring and a double-buffered TMEM accumulator, so prologue/epilogue/TMEM
//     allocation are amortised over all tiles of the launch.
#include <cuda.h>

#include <cstdlib>

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
constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpilogueThreads = 128;
constexpr int kThreads = kProducerThreads + kEpilogueThreads + 32;     // 416
constexpr int kRowsPerThread = BLOCK_M * 8 / kProducerThreads;          // 4

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}
// relu(x*s + b) on two packed bf16 values

template <int BLOCK_N, bool TEPI = false> struct Cfg {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    // TEPI (TMA epilogue, see conv_gemm.cu): one bf16 output tile is staged in shared memory
    static constexpr int kStgBytes = TEPI ? BLOCK_M * BLOCK_N * 2 : 0;
    static constexpr int kStages = (BLOCK_N >= 256) ? (TEPI ? 3 : 4) : (BLOCK_N >= 128 ? 5 : 6);
    static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : 2 * BLOCK_N;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStgBytes + 1024 + 256;
};

// debug aid: per-stage role timestamps of CTA 0 (SPK_TC2_DBG=1), read back by spk_debug_tc2_timeline
__device__ long long g_tc2_ts[256 * 8];
__device__ int g_tc2_dbg;
#define TC2_TS(idx, slot) do { if (dbg && (idx) < 256) g_tc2_ts[(idx) * 8 + (slot)] = clock64(); } while (0)

struct GatherRegs {
    uint4 v[kRowsPerThread];
    uint32_t pad_mask;     // bit i: piece i is zero padding (prologue must not touch it)
};

template <int BLOCK_N, typename TOut, typename TRes, bool TEPI>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const ConvArgs a, const bf16 *__restrict__ pro_scale_bf, const bf16 *__restrict__ pro_shift_bf,
                int n_tiles_n, long long n_tiles, const __grid_constant__ CUtensorMap wmap,
                const __grid_constant__ CUtensorMap ymap, const __grid_constant__ CUtensorMap rmap) {
    using C = Cfg<BLOCK_N, TEPI>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t s_stg = base + C::kStages * C::kStageBytes;       // TEPI: output staging tile
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (base - raw) + C::kStages * C::kStageBytes + C::kStgBytes);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
    auto accf_bar = [&](int b) { return bar0 + 8u * (2 * C::kStages + b); };
    auto acce_bar = [&](int b) { return bar0 + 8u * (2 * C::kStages + 2 + b); };
    auto rfull_bar = [&]() { return bar0 + 8u * (2 * C::kStages + 4); };     // TEPI: residual tile landed in the staging buffer
    auto sfree_bar = [&]() { return bar0 + 8u * (2 * C::kStages + 5); };     // TEPI: the TMA store has read the staging buffer
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bars + 2 * C::kStages + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kMmaWarp = kProducerWarps + 4;      // warps 0-7 producers, 8-11 epilogue, 12 MMA
    const bool dbg = g_tc2_dbg != 0 && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kMmaWarp || warp == kProducerWarps);
    int dbg_idx = 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(full_bar(s), kProducerWarps + 1);       // one arrival per gather warp + the TMA expect_tx arrive
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accf_bar(b), 1);
            mbar_init(acce_bar(b), kEpilogueThreads);
        }
        mbar_init(rfull_bar(), 1);
        mbar_init(sfree_bar(), 1);
        fence_barrier_init();
        prefetch_tmap(&wmap);
        if (TEPI) {
            prefetch_tmap(&ymap);
            prefetch_tmap(&rmap);
        }
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nk = (a.K + BLOCK_K - 1) / BLOCK_K;
    const int HoWo = a.Ho * a.Wo;

    if (warp < kProducerWarps) {
        // =========================== producers ===========================
        const int t = threadIdx.x;                 // 0..255
        const int j = t & 7;                       // 16-byte column of the 128-byte row
        const int r0 = t >> 3;                     // rows r0 + 32*i
        const uint32_t sw_off = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
        const bf16 *x = static_cast<const bf16 *>(a.x);
        const bool has_pro = pro_scale_bf != nullptr;

        // Decoded rows of the tile whose loads are being issued.  The gather warps are instruction-latency bound (two
        // warps per scheduler, one dependent chain each), so everything that does not change from chunk to chunk is
        // hoisted: per row the element offset of its (kh, kw) = (0, 0) pixel and two bit masks saying which kh / kw
        // stay inside the image; per thread the (kh, kw, channel) position of its 16-byte column, advanced by 64
        // channels per chunk without divisions.  A chunk then costs two shifts, an AND and one add per row.
        long long base0[kRowsPerThread];
        uint32_t hmask[kRowsPerThread], wmask[kRowsPerThread];
        int pix0[kRowsPerThread], hi0[kRowsPerThread], wi0[kRowsPerThread];       // reflect / large-kernel path only
        const bool masked = !a.pad_reflect && a.KH <= 32 && a.KW <= 32;
        auto decode = [&](long long tile) {
            const long long m0 = (tile / n_tiles_n) * BLOCK_M;
#pragma unroll
            for (int i = 0; i < kRowsPerThread; ++i) {
                const long long m = m0 + r0 + 32 * i;
                hmask[i] = 0u;
                wmask[i] = 0u;
                base0[i] = 0;
                if (m < a.M) {
                    const int b = (int)(m / HoWo);
                    const int r = (int)(m - (long long)b * HoWo);
                    const int ho = r / a.Wo, wo = r - ho * a.Wo;
                    pix0[i] = b * a.H * a.W;
                    hi0[i] = ho * a.sh - a.ph;
                    wi0[i] = wo * a.sw - a.pw;
                    if (masked) {
                        for (int kh = 0; kh < a.KH; ++kh) {
                            const int hi = hi0[i] + kh * a.dh;
                            if (hi >= 0 && hi < a.H) hmask[i] |= 1u << kh;
                        }
                        for (int kw = 0; kw < a.KW; ++kw) {
                            const int wi = wi0[i] + kw * a.dw;
                            if (wi >= 0 && wi < a.W) wmask[i] |= 1u << kw;
                        }
                        base0[i] = ((long long)pix0[i] + (long long)hi0[i] * a.W + wi0[i]) * a.in_ld + a.in_choff;
                    }
                } else {
                    pix0[i] = 0;
                    hi0[i] = -(1 << 28);
                    wi0[i] = 0;
                }
            }
        };
        // (kh, kw, c) of this thread's column in the chunk being issued
        int ckh = 0, ckw = 0, cc = 0;
        auto col_reset = [&]() {
            ckh = 0; ckw = 0; cc = j * 8;
            while (cc >= a.Cin) { cc -= a.Cin; if (++ckw == a.KW) { ckw = 0; ++ckh; } }
        };
        auto col_advance = [&]() {
            cc += BLOCK_K;
            while (cc >= a.Cin) { cc -= a.Cin; if (++ckw == a.KW) { ckw = 0; ++ckh; } }
        };
        auto issue = [&](int kc, GatherRegs &g) {
            g.pad_mask = 0;
            if (kc == 0) col_reset();
            const bool k_ok = ckh < a.KH;
            if (masked) {
                const long long delta = ((long long)(ckh * a.dh) * a.W + ckw * a.dw) * a.in_ld + cc;
#pragma unroll
                for (int i = 0; i < kRowsPerThread; ++i) {
                    if (k_ok && ((hmask[i] >> ckh) & (wmask[i] >> ckw) & 1u)) {
                        g.v[i] = ldg16(x + base0[i] + delta);
                    } else {
                        g.v[i] = make_uint4(0u, 0u, 0u, 0u);
                        g.pad_mask |= 1u << i;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < kRowsPerThread; ++i) {
                    const int hi = hi0[i] + ckh * a.dh;
                    int wi = wi0[i] + ckw * a.dw;
                    if (a.pad_reflect) wi = wi < 0 ? -wi : (wi >= a.W ? 2 * (a.W - 1) - wi : wi);
                    if (k_ok && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) {
                        const long long off = ((long long)pix0[i] + (long long)hi * a.W + wi) * a.in_ld + a.in_choff + cc;
                        g.v[i] = ldg16(x + off);
                    } else {
                        g.v[i] = make_uint4(0u, 0u, 0u, 0u);
                        g.pad_mask |= 1u << i;
                    }
                }
            }
            col_advance();
        };
        int stage = 0;
        uint32_t phase = 0;
        // finish chunk kc of 'tile': prologue, wait for the slot, TMA the weights, store A, signal
        auto finish = [&](long long tile, int kc, GatherRegs &g) {
            if (has_pro) {
                const int k = kc * BLOCK_K + j * 8;
                if (k < a.K) {
                    const int c = k % a.Cin;
                    const uint4 s4 = ldg16(pro_scale_bf + c), h4 = ldg16(pro_shift_bf + c);
                    const bool relu = a.pro_relu != 0;
#pragma unroll
                    for (int i = 0; i < kRowsPerThread; ++i) {
                        if (g.pad_mask & (1u << i)) continue;      // zero padding stays zero
                        g.v[i].x = bnrelu2(g.v[i].x, s4.x, h4.x, relu);
                        g.v[i].y = bnrelu2(g.v[i].y, s4.y, h4.y, relu);
                        g.v[i].z = bnrelu2(g.v[i].z, s4.z, h4.z, relu);
                        g.v[i].w = bnrelu2(g.v[i].w, s4.w, h4.w, relu);
                    }
                }
            }
            TC2_TS(dbg_idx, 0);
            mbar_wait(empty_bar(stage), phase ^ 1u);
            TC2_TS(dbg_idx, 1);
            const uint32_t sa = base + stage * C::kStageBytes;
            if (t == 0) {
                const int nt = (int)(tile % n_tiles_n);
                mbar_arrive_expect_tx(full_bar(stage), C::kBBytes);
                tma_load_2d(sa + C::kABytes, &wmap, kc * BLOCK_K, nt * BLOCK_N, full_bar(stage));
            }
#pragma unroll
            for (int i = 0; i < kRowsPerThread; ++i) sts16(sa + sw_off + i * 4096, g.v[i]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(stage));        // 8 arrivals per stage instead of 256 on one barrier word
            TC2_TS(dbg_idx, 2);
            ++dbg_idx;
            if (++stage == C::kStages) {
                stage = 0;
                phase ^= 1u;
            }
        };

        long long itile = blockIdx.x, ftile = blockIdx.x;      // issue cursor, finish cursor
        int ikc = 0, fkc = 0;
        bool decoded = false;
        if (!has_pro && masked) {
            // ---- asynchronous gather: the 16-byte pieces go straight to the stage's swizzled slot with cp.async (zero
            // padding: a plain store), kAhead chunks ahead of the one being completed; "complete" = this thread's copy
            // group of that chunk has landed (cp.async.wait_group), then proxy fence + one arrival per warp.  Register
            // staging cannot run this deep: the loads of all chunks in flight share a few counting scoreboards, so
            // waiting for the oldest chunk waited for the newest one too (measured: 2 -> 4 register sets, no change).
            constexpr int kAhead = C::kStages > 3 ? 3 : C::kStages - 1;
            static_assert(kAhead < C::kStages, "the stage ring must hold the chunks in flight");
            int istage = 0;
            uint32_t iphase = 0;
            auto issue_async = [&]() {
                if (itile < n_tiles) {
                    if (!decoded) { decode(itile); decoded = true; }
                    if (ikc == 0) col_reset();
                    const bool k_ok = ckh < a.KH;
                    const long long delta = ((long long)(ckh * a.dh) * a.W + ckw * a.dw) * a.in_ld + cc;
                    mbar_wait(empty_bar(istage), iphase ^ 1u);
                    const uint32_t sa = base + istage * C::kStageBytes;
                    if (t == 0) {
                        const int nt = (int)(itile % n_tiles_n);
                        mbar_arrive_expect_tx(full_bar(istage), C::kBBytes);
                        tma_load_2d(sa + C::kABytes, &wmap, ikc * BLOCK_K, nt * BLOCK_N, full_bar(istage));
                    }
#pragma unroll
                    for (int i = 0; i < kRowsPerThread; ++i) {
                        const uint32_t dst = sa + sw_off + i * 4096;
                        if (k_ok && ((hmask[i] >> ckh) & (wmask[i] >> ckw) & 1u))
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(x + base0[i] + delta) : "memory");
                        else
                            sts16(dst, make_uint4(0u, 0u, 0u, 0u));
                    }
                    col_advance();
                    if (++ikc == nk) { ikc = 0; itile += gridDim.x; decoded = false; }
                    if (++istage == C::kStages) { istage = 0; iphase ^= 1u; }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");      // always: keeps the group count uniform at the tail
            };
#pragma unroll
            for (int d = 0; d < kAhead; ++d) issue_async();
            while (ftile < n_tiles) {
                issue_async();
                asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory");
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(stage));
                if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                if (++fkc == nk) { fkc = 0; ftile += gridDim.x; }
            }
        } else {
        // ---- register-staged gather (BN-ReLU prologue, reflect padding): one chunk ahead
        constexpr int kDepth = 2;
        GatherRegs g[kDepth];
        auto issue_next = [&](GatherRegs &dst) {
            if (itile >= n_tiles) return;
            if (!decoded) { decode(itile); decoded = true; }
            issue(ikc, dst);
            if (++ikc == nk) { ikc = 0; itile += gridDim.x; decoded = false; }
        };
#pragma unroll
        for (int d = 0; d < kDepth - 1; ++d) issue_next(g[d]);
        while (ftile < n_tiles) {
#pragma unroll
            for (int s = 0; s < kDepth; ++s) {
                if (ftile < n_tiles) {
                    issue_next(g[(s + kDepth - 1) % kDepth]);
                    finish(ftile, fkc, g[s]);
                    if (++fkc == nk) { fkc = 0; ftile += gridDim.x; }
                }
            }
        }
        }
    } else if (warp == kMmaWarp) {
        // =========================== MMA issuer ===========================
        constexpr uint32_t idesc = make_idesc(BLOCK_N);
        int stage = 0;
        uint32_t phase = 0, it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, acc_phase = (it >> 1) & 1u;
            mbar_wait(acce_bar(buf), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
            for (int kc = 0; kc < nk; ++kc) {
                TC2_TS(dbg_idx, 3);
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                TC2_TS(dbg_idx, 4);
                if (lane == 0) {
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint64_t ad = make_desc_sw128(sa), bd = make_desc_sw128(sa + C::kABytes);
#pragma unroll
                    for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                        umma_bf16(d_tmem, ad + 2u * kk, bd + 2u * kk, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(empty_bar(stage));
                    if (kc == nk - 1) umma_commit(accf_bar(buf));
                }
                __syncwarp();
                TC2_TS(dbg_idx, 5);
                ++dbg_idx;
                if (++stage == C::kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;                       // warps 8..11 -> TMEM lane quarters 0..3
        const int row = q * 32 + lane;
        TOut *y = static_cast<TOut *>(a.y);
        const TRes *res = static_cast<const TRes *>(a.res);
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
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
                // ---- TMA epilogue (same scheme as conv_gemm.cu): tile -> 128B-swizzled staging blocks of 64 columns
                // -> TMA store; a residual tile is TMA-loaded into the staging buffer by the elected epilogue thread
                // (after the previous store has read it) and updated in place
                if (it == 0 && res != nullptr && warp == kProducerWarps && elect_one()) {
                    const int nblk = min(BLOCK_N / 64, (a.Cout - n0 + 63) / 64);
                    mbar_arrive_expect_tx(rfull_bar(), (uint32_t)nblk * (BLOCK_M * 128u));
                    for (int j = 0; j < nblk; ++j)
                        tma_load_2d(s_stg + (uint32_t)j * (BLOCK_M * 128u), &rmap, n0 + 64 * j, (int)(mt * BLOCK_M), rfull_bar());
                }
                if (res != nullptr) mbar_wait(rfull_bar(), it & 1u);
                else mbar_wait(sfree_bar(), (it & 1u) ^ 1u);
                mbar_wait(accf_bar(buf), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
                const uint32_t srow = s_stg + (uint32_t)row * 128u, x7 = (uint32_t)(row & 7);
#pragma unroll 1
                for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
                    const int n = n0 + c0;
                    if (n >= a.Cout) break;
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    float4 s4[4], h4[4];
                    if (a.epi_scale != nullptr) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            s4[e] = __ldg(reinterpret_cast<const float4 *>(a.epi_scale + n + 4 * e));
                            h4[e] = __ldg(reinterpret_cast<const float4 *>(a.epi_shift + n + 4 * e));
                        }
                    }
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
                    if (a.epi_scale != nullptr) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            v[4 * e] = fmaf(v[4 * e], s4[e].x, h4[e].x); v[4 * e + 1] = fmaf(v[4 * e + 1], s4[e].y, h4[e].y);
                            v[4 * e + 2] = fmaf(v[4 * e + 2], s4[e].z, h4[e].z); v[4 * e + 3] = fmaf(v[4 * e + 3], s4[e].w, h4[e].w);
                        }
                    }
                    const uint32_t blk = srow + (uint32_t)(c0 >> 6) * (BLOCK_M * 128u);
                    const uint32_t ch0 = (uint32_t)((c0 & 63) >> 3);
                    if (res != nullptr) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
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
                        for (int e = 0; e < 16; e += 4) {
                            const float4 p4 = __ldg(reinterpret_cast<const float4 *>(a.post_scale + n + e));
                            const float4 q4 = __ldg(reinterpret_cast<const float4 *>(a.post_shift + n + e));
                            v[e] = fmaf(v[e], p4.x, q4.x); v[e + 1] = fmaf(v[e + 1], p4.y, q4.y);
                            v[e + 2] = fmaf(v[e + 2], p4.z, q4.z); v[e + 3] = fmaf(v[e + 3], p4.w, q4.w);
                        }
                        apply_act_vec(v, a.post_act);
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        sts16(blk + (((ch0 + e) ^ x7) << 4),
                              make_uint4(pack2(v[8 * e], v[8 * e + 1]), pack2(v[8 * e + 2], v[8 * e + 3]), pack2(v[8 * e + 4], v[8 * e + 5]),
                                         pack2(v[8 * e + 6], v[8 * e + 7])));
                }
                tc_fence_before();
                mbar_arrive(acce_bar(buf));
                fence_proxy_async();
                epi_bar_sync();
                if (warp == kProducerWarps && elect_one()) {
                    const int nblk = min(BLOCK_N / 64, (a.Cout - n0 + 63) / 64);
                    for (int j = 0; j < nblk; ++j)
                        tma_store_2d(&ymap, n0 + 64 * j, (int)(mt * BLOCK_M), s_stg + (uint32_t)j * (BLOCK_M * 128u));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    const long long ntile = tile + gridDim.x;
                    if (res != nullptr) {
                        if (ntile < n_tiles) {          // residual of the next tile into the (now free) staging buffer
                            const long long nmt = ntile / n_tiles_n;
                            const int nn0 = (int)(ntile - nmt * n_tiles_n) * BLOCK_N;
                            const int nb2 = min(BLOCK_N / 64, (a.Cout - nn0 + 63) / 64);
                            mbar_arrive_expect_tx(rfull_bar(), (uint32_t)nb2 * (BLOCK_M * 128u));
                            for (int j = 0; j < nb2; ++j)
                                tma_load_2d(s_stg + (uint32_t)j * (BLOCK_M * 128u), &rmap, nn0 + 64 * j, (int)(nmt * BLOCK_M), rfull_bar());
                        }
                    } else {
                        mbar_arrive(sfree_bar());
                    }
                }
                __syncwarp();
                continue;
            }
            mbar_wait(accf_bar(buf), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
                const int n = n0 + c0;
                if (n >= a.Cout) break;               // uniform across the CTA
                uint32_t r[16];
                tmem_ld16(taddr + c0, r);
                tmem_ld_wait();
                if (m < a.M) {
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
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
                        const TRes *rp = res + m * a.res_ld + a.res_choff + n;
                        if constexpr (sizeof(TRes) == 2) {       // 16-byte loads (res_ld, res_choff are multiples of 8)
#pragma unroll
                            for (int e = 0; e < 16; e += 8) {
                                const uint4 t = ldg16(rp + e);
                                const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                                for (int h = 0; h < 4; ++h) {
                                    const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162 *>(&w4[h]);
                                    v[e + 2 * h] += __low2float(b2);
                                    v[e + 2 * h + 1] += __high2float(b2);
                                }
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; e += 4) {
                                const float4 t = __ldg(reinterpret_cast<const float4 *>(rp + e));
                                v[e] += t.x; v[e + 1] += t.y; v[e + 2] += t.z; v[e + 3] += t.w;
                            }
                        }
                    }
                    if (grow != nullptr && a.gate_additive) {
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
                    TOut *yp = y + m * a.out_ld + a.out_choff + n;
                    if constexpr (sizeof(TOut) == 2) {
                        *reinterpret_cast<uint4 *>(yp) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                        *reinterpret_cast<uint4 *>(yp + 8) = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            *reinterpret_cast<float4 *>(yp + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acce_bar(buf));
        }
        if (TEPI) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the last stores have reached global memory
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------ host: tensor maps
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

struct MapKey {
    const void *w; int K, Cout, block_n;
    bool operator<(const MapKey &o) const {
        if (w != o.w) return w < o.w;
        if (K != o.K) return K < o.K;
        if (Cout != o.Cout) return Cout < o.Cout;
        return block_n < o.block_n;
    }
};

int weight_map(const void *w, int K, int Cout, int block_n, CUtensorMap *out) {
    static std::mutex mu;
    static std::map<MapKey, CUtensorMap> cache;
    std::lock_guard<std::mutex> lk(mu);
    MapKey key{w, K, Cout, block_n};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return SPK_OK;
    }
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SPK_ERR_CUDA;
    }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(bf16)};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for K=%d Cout=%d box_n=%d", (int)r, K, Cout, block_n);
        return SPK_ERR_CUDA;
    }
    cache[key] = m;
    *out = m;
    return SPK_OK;
}

// bf16 copies of prologue scale/shift vectors, keyed by the fp32 device pointer
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
    // +8 elements of slack: the producer reads whole 16-byte pieces
    SPK_CUDA_OK(cudaMalloc(&d, (size_t)(n + 8) * sizeof(bf16)));
    SPK_CUDA_OK(cudaMemsetAsync(d, 0, (size_t)(n + 8) * sizeof(bf16), s));
    int rc = launch_f32_to_bf16(src, d, n, s);
    if (rc != SPK_OK) return rc;
    cache[src] = d;
    *out = d;
    return SPK_OK;
}

// [rows, cols] bf16 matrix with row pitch ld: boxes of 64 columns x 128 rows, 128B swizzle (epilogue staging blocks)
int tile_map(const void *ptr, long long rows, int cols, long long ld, CUtensorMap *out) {
    const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)ld * 2};
    const uint32_t box[2] = {64u, (uint32_t)BLOCK_M};
    return tmap_encode_bf16(ptr, 2, dims, strides, box, 128, out);
}

template <int BLOCK_N, typename TOut, typename TRes, bool TEPI = false>
int launch_one(const ConvArgs &a, cudaStream_t s) {
    using C = Cfg<BLOCK_N, TEPI>;
    auto kern = conv_tc2_kernel<BLOCK_N, TOut, TRes, TEPI>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes); });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_tc2) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    CUtensorMap wmap;
    int rc = weight_map(a.w, a.K, a.Cout, BLOCK_N, &wmap);
    if (rc != SPK_OK) return rc;
    const bf16 *ps = nullptr, *ph = nullptr;
    if (a.pro_scale != nullptr) {
        if (a.pro_scale_bf != nullptr && a.pro_shift_bf != nullptr) {      // model-owned copies made at set_program time
            ps = static_cast<const bf16 *>(a.pro_scale_bf);
            ph = static_cast<const bf16 *>(a.pro_shift_bf);
        } else {
            rc = bf16_vector(a.pro_scale, a.Cin, &ps, s);
            if (rc == SPK_OK) rc = bf16_vector(a.pro_shift, a.Cin, &ph, s);
            if (rc != SPK_OK) return rc;
        }
    }
    CUtensorMap ymap = wmap, rmap = wmap;        // placeholders unless the TMA epilogue is used
    if (TEPI) {
        rc = tile_map(static_cast<const bf16 *>(a.y) + a.out_choff, a.M, a.Cout, a.out_ld, &ymap);
        if (rc == SPK_OK && a.res != nullptr) rc = tile_map(static_cast<const bf16 *>(a.res) + a.res_choff, a.M, a.Cout, a.res_ld, &rmap);
        if (rc != SPK_OK) return rc;
    }
    const long long mt = (a.M + BLOCK_M - 1) / BLOCK_M;
    const int ntn = (a.Cout + BLOCK_N - 1) / BLOCK_N;
    const long long tiles = mt * ntn;
    long long grid = std::min<long long>(tiles, sm_count());
    kern<<<(unsigned)grid, kThreads, C::kSmemBytes, s>>>(a, ps, ph, ntn, tiles, wmap, ymap, rmap);
    return check_launch("conv_tc2_kernel");
}

template <typename TOut, typename TRes>
int launch_n(const ConvArgs &a, cudaStream_t s) {
    if (a.Cout <= 32) return launch_one<32, TOut, TRes>(a, s);
    if constexpr (sizeof(TOut) == 2 && sizeof(TRes) == 2) {
        // bf16 tiles leave through the TMA epilogue (SPK_TC2_TEPI=0 selects the register epilogue for A/B runs)
        static const bool off = [] { const char *e = getenv("SPK_TC2_TEPI"); return e && e[0] == '0'; }();
        const bool ok = a.gate == nullptr && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0 &&
                        (a.res == nullptr || (reinterpret_cast<uintptr_t>(a.res) & 15) == 0);
        if (ok && !off) {
            if (a.Cout <= 64) return launch_one<64, TOut, TRes, true>(a, s);
            if (a.Cout <= 128) return launch_one<128, TOut, TRes, true>(a, s);
            return launch_one<256, TOut, TRes, true>(a, s);
        }
    }
    if (a.Cout <= 64) return launch_one<64, TOut, TRes>(a, s);
    if (a.Cout <= 128) return launch_one<128, TOut, TRes>(a, s);
    return launch_one<256, TOut, TRes>(a, s);
}

}  // namespace

int launch_conv_tc2(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s) {
    if (a.M == 0) return SPK_OK;
    const bool res_bf16 = a.res == nullptr ? (out_dtype == SPK_DT_BF16) : (res_dtype == SPK_DT_BF16);
    if (out_dtype == SPK_DT_BF16) {
        if (!res_bf16) {
            set_error("conv_tc2: a bf16 output takes a bf16 residual");
            return SPK_ERR_UNSUPPORTED;
        }
        return launch_n<bf16, bf16>(a, s);
    }
    return res_bf16 ? launch_n<float, bf16>(a, s) : launch_n<float, float>(a, s);
}


// ---- entry points of the generic tcgen05 path (the first-generation gather kernel these used to select between
// is gone: this kernel superseded it in every shape)
bool conv_tc_supported(const ConvArgs &a, int in_dtype) {
    if (in_dtype != SPK_DT_BF16) return false;
    if (a.Cin % 8 != 0 || a.in_ld % 8 != 0 || a.in_choff % 8 != 0) return false;
    if (a.Cout % 16 != 0 || a.out_ld % 8 != 0 || a.out_choff % 8 != 0) return false;
    if (a.res != nullptr && (a.res_ld % 8 != 0 || a.res_choff % 8 != 0)) return false;
    if (a.M < BLOCK_M) return false;          // tiny problems (e.g. per-segment dense) stay on CUDA cores
    return true;
}

int launch_conv_tc(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s) {
    return launch_conv_tc2(a, out_dtype, res_dtype, s);
}

}  // namespace spk

// debug aids (not part of the ABI)
extern "C" int spk_debug_tc2_enable(int on) { return cudaMemcpyToSymbol(spk::g_tc2_dbg, &on, sizeof(int)) == cudaSuccess ? 0 : -1; }
extern "C" int spk_debug_tc2_timeline(long long *dst) {
    return cudaMemcpyFromSymbol(dst, spk::g_tc2_ts, sizeof(long long) * 256 * 8) == cudaSuccess ? 0 : -1;
}
