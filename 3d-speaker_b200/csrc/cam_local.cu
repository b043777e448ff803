// Fused CAM layer for sm_100a: the dilated k=3 "local" conv of CAMLayer, the context gate and the
// gating multiply in ONE kernel (speakerlab/models/campplus/layers.py:93-110):
//
//   y   = linear_local(x)                         Conv1d(128 -> 32, k=3, dilation d, padding d)
//   ctx = mean_T(x) + seg_avg_100(x)              full-utterance mean + 100-frame segment means
//   m   = sigmoid(W2 relu(W1 ctx + b1) + b2)      1x1 convs with bias, evaluated once per window
//   out = y * m                                   written into the block's concat buffer slice
//
// A CTA item is a group of whole segments.  Their rows (zero padded by d on both sides, pitch P)
// are staged ONCE in shared memory by TMA: one 3-D box {64 channels, P frames, 1 segment} per
// segment and channel half, starting at frame -d, so the padding rows - and the rows of segments past
// the end of the batch - are the TMA unit's out-of-bounds zero fill.  The slab is two K-major
// 128-byte-swizzled halves (channels 0-63 and 64-127).  From that one copy
//   * the MMA warp runs the three taps as row-shifted UMMA descriptors (tap k = start address +
//     k*d rows; the swizzle is a function of the absolute shared-memory address, so any row shift
//     keeps the pattern), tcgen05.mma M=128 N=32 K=16, accumulators in TMEM;
//   * the same warp gets the per-window column sums from the tensor core too: the slab read as an
//     MN-major operand (M = channels, K = rows) times a constant 0/1 window indicator matrix gives
//     sum_t x[t, c] per (segment, window) in fp32, in a fixed order;
//   * four "gate" warps read those sums from TMEM (one channel per thread) and run the two tiny
//     mat-vecs of the gate while the conv MMAs are in flight;
//   * four epilogue warps multiply the gate into the accumulator on its way to HBM.
// Compared with the unfused path this removes one kernel launch, the separate read of x for the
// context and the 3x gather of x for the taps.  Slabs and TMEM accumulators are double buffered,
// so staging, MMA and epilogue of consecutive items overlap.  Small-N MMAs are bound by the
// shared-memory read of A (4 KB per MMA, ~40 cycles) and by the issue rate of the one issuing
// thread, hence the fully unrolled issue loops.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "ops.cuh"
#include "tc.cuh"
#include "tmap.cuh"

namespace spk {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kCin = 128, kCout = 32, kPlanes = kCin / 8, kTaps = 3;
constexpr int kGate = 128, kEpi = 128, kThreads = 32 + 32 + kGate + kEpi;       // 320: TMA warp, MMA warp, gate, epilogue
constexpr int kMaxSeg = 8, kMaxWin = 4, kMaxHidden = 64;
// mbarrier / TMEM / tcgen05 / descriptor wrappers: tc.cuh (shared with the slab and GEMM kernels)
using namespace tc;
__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    umma_bf16_acc(tmem_d, adesc, bdesc, idesc);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return tc::desc_hi(sbo_bytes, kLayoutNone); }
__device__ __forceinline__ uint32_t desc_hi_sw128(uint32_t sbo_bytes) { return tc::desc_hi(sbo_bytes, kLayoutSw128); }
// sums: D[128 channels][16] = A^T-view (MN-major, bit 15) x indicator (K-major)
constexpr int kSumCols = 16;
constexpr uint32_t kIdescSum = idesc_bf16(kSumCols, 1);
constexpr uint32_t kIdescN32 = idesc_bf16(32);
__device__ __forceinline__ void gate_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// debug aid: per-item role timestamps of CTA 0 (SPK_CAM_DBG & 64), read back by spk_debug_cam_timeline
__device__ long long g_cam_ts[16 * 12];
#define CAM_TS(slot) do { if ((dbg & 64) && blockIdx.x == 0 && it < 16 && (threadIdx.x & 31) == 0) g_cam_ts[it * 12 + (slot)] = clock64(); } while (0)

// The two mat-vecs of the gate for NC (segment, window) contexts at once, on 4 warps, with one block barrier.
// Both layers split the outputs across warps and the reduction dimension across the four lane groups
// ks = lane >> 3 of a warp, so partial sums are folded with two shuffles (fixed order) instead of shared memory:
//   hidden = relu(W1 ctx + b1): warp w owns hidden units 16w..16w+15, lane = (pair jq = lane & 7, ks);
//                               lane group ks covers channels ch = 4i + ks
//   gate   = sigmoid(W2 hidden + b2): warp w owns outputs 8w..8w+7, lane = (o = lane & 7, ks);
//                               lane group ks covers hidden units 4i + ks
// ctx and hid are stored permuted ([cb][ks][i]) so a lane reads its operands as float4.  A thread needs the same
// 64 + 16 weights for every item, so they live in REGISTERS for the life of the CTA (loaded once from the transposed
// global copies): the gate warps run while the MMA warp saturates the shared-memory read port, and weight reads from
// shared memory were what made this phase the slowest stage of the pipeline.
template <int NC>
__device__ __forceinline__ void gate_mlp(const float *ctx, float *hid, float *gateb, const float2 (&w1r)[32], const float (&w2r)[16],
                                         float2 b1r, float b2r, int hidden, int t) {
    const int w = t >> 5, lane = t & 31, ks = lane >> 3;
    {
        const int j = 16 * w + 2 * (lane & 7);
        float2 acc[NC];
#pragma unroll
        for (int cb = 0; cb < NC; ++cb) acc[cb] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
#pragma unroll
            for (int cb = 0; cb < NC; ++cb) {
                const float4 x = *reinterpret_cast<const float4 *>(&ctx[cb * kCin + ks * 32 + i]);
                acc[cb] = __ffma2_rn(w1r[i], make_float2(x.x, x.x), acc[cb]);
                acc[cb] = __ffma2_rn(w1r[i + 1], make_float2(x.y, x.y), acc[cb]);
                acc[cb] = __ffma2_rn(w1r[i + 2], make_float2(x.z, x.z), acc[cb]);
                acc[cb] = __ffma2_rn(w1r[i + 3], make_float2(x.w, x.w), acc[cb]);
            }
        }
#pragma unroll
        for (int cb = 0; cb < NC; ++cb) {
            acc[cb].x += __shfl_xor_sync(0xffffffffu, acc[cb].x, 8);
            acc[cb].y += __shfl_xor_sync(0xffffffffu, acc[cb].y, 8);
            acc[cb].x += __shfl_xor_sync(0xffffffffu, acc[cb].x, 16);
            acc[cb].y += __shfl_xor_sync(0xffffffffu, acc[cb].y, 16);
        }
        if (ks == 0) {
#pragma unroll
            for (int cb = 0; cb < NC; ++cb) {
                // unit j lives at [cb][j & 3][j >> 2]
                hid[cb * kMaxHidden + (j & 3) * 16 + (j >> 2)] = j < hidden ? fmaxf(acc[cb].x + b1r.x, 0.f) : 0.f;
                hid[cb * kMaxHidden + ((j + 1) & 3) * 16 + ((j + 1) >> 2)] = j + 1 < hidden ? fmaxf(acc[cb].y + b1r.y, 0.f) : 0.f;
            }
        }
    }
    gate_bar_sync();
    {
        const int o = 8 * w + (lane & 7);
        float acc[NC];
#pragma unroll
        for (int cb = 0; cb < NC; ++cb) acc[cb] = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
#pragma unroll
            for (int cb = 0; cb < NC; ++cb) {
                const float4 h = *reinterpret_cast<const float4 *>(&hid[cb * kMaxHidden + ks * 16 + i]);
                acc[cb] = fmaf(w2r[i], h.x, acc[cb]);
                acc[cb] = fmaf(w2r[i + 1], h.y, acc[cb]);
                acc[cb] = fmaf(w2r[i + 2], h.z, acc[cb]);
                acc[cb] = fmaf(w2r[i + 3], h.w, acc[cb]);
            }
        }
#pragma unroll
        for (int cb = 0; cb < NC; ++cb) {
            acc[cb] += __shfl_xor_sync(0xffffffffu, acc[cb], 8);
            acc[cb] += __shfl_xor_sync(0xffffffffu, acc[cb], 16);
        }
        if (ks == 0) {
#pragma unroll
            for (int cb = 0; cb < NC; ++cb) gateb[cb * kCout + o] = 1.f / (1.f + expf(-(acc[cb] + b2r)));
        }
    }
}

// ctx = mean over the segment + mean over the window for one channel, from the raw window sums (registers only).
// Written permuted: channel ch of combo cb lives at ctx[cb][ch & 3][ch >> 2].
template <int NWIN>
__device__ __forceinline__ void ctx_from_sums(const float (&sums)[8], float *ctx, int G, int ch, float inv_T, float inv_seg,
                                              float inv_last) {
    const int pos = (ch & 3) * 32 + (ch >> 2);
#pragma unroll
    for (int gs = 0; gs < 8 / NWIN; ++gs) {
        if (gs < G) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < NWIN; ++w) tot += sums[gs * NWIN + w];
            tot *= inv_T;
#pragma unroll
            for (int w = 0; w < NWIN; ++w)
                ctx[(gs * NWIN + w) * kCin + pos] = fmaf(sums[gs * NWIN + w], w == NWIN - 1 ? inv_last : inv_seg, tot);
        }
    }
}

struct CamGeom {
    int T, d, P, G, n_tiles, px, nwin, seg_len, hidden;
    int smem_bytes, tmem_cols;
    unsigned p_magic;        // ceil(2^32 / P)
    uint32_t off_w, off_slab, slab_bytes, half_bytes, off_part, off_win, off_ind, off_hid, off_gate, off_bar, off_mlp;
    int k16;                 // 16-row K steps of the column-sum MMAs (covers G*P rows)
};

__global__ void __launch_bounds__(kThreads, 1)
cam_local_kernel(const ConvArgs a, const CamGeom g, const float *__restrict__ w1t, const float *__restrict__ b1,
                 const float *__restrict__ w2t, const float *__restrict__ b2, int n_items, int dbg,
                 const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t s0 = smem_u32(smem);
    const uint32_t s_w = s0 + g.off_w, s_slab0 = s0 + g.off_slab, s_bar = s0 + g.off_bar, s_ind = s0 + g.off_ind;
    float *win = reinterpret_cast<float *>(smem + g.off_win);       // [G*nwin][128]  the contexts
    float *hid = reinterpret_cast<float *>(smem + g.off_hid);       // [G*nwin][64]
    float *gate = reinterpret_cast<float *>(smem + g.off_gate);     // [2][G][nwin][32]
    auto sfull = [&](int i) { return s_bar + 8u * i; };             // slab staged           (TMA tx -> MMA)
    auto sempty = [&](int i) { return s_bar + 8u * (2 + i); };      // slab consumed         (MMA commit -> TMA warp)
    auto afull = [&](int i) { return s_bar + 8u * (4 + i); };       // conv accumulator done (MMA commit -> epilogue)
    auto aempty = [&](int i) { return s_bar + 8u * (6 + i); };      // TMEM buffer drained   (epilogue -> MMA)
    auto gfull = [&](int i) { return s_bar + 8u * (8 + i); };       // gate values ready     (gate -> epilogue)
    auto gempty = [&](int i) { return s_bar + 8u * (10 + i); };     // gate values consumed  (epilogue -> gate)
    auto cfull = [&](int i) { return s_bar + 8u * (12 + i); };      // column sums done      (MMA commit -> gate)
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + g.off_bar + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t acc_cols = (uint32_t)g.n_tiles * 32u + 2 * kSumCols;  // per TMEM buffer: conv tiles, then the two sum halves

    pdl_trigger();
    if (threadIdx.x == 0) {
        tmap_prefetch(&xmap);
        for (int i = 0; i < 2; ++i) {
            mbar_init(sfull(i), 1);
            mbar_init(sempty(i), 1);
            mbar_init(afull(i), 1);
            mbar_init(aempty(i), kEpi);
            mbar_init(gfull(i), kGate);
            mbar_init(gempty(i), kEpi);
            mbar_init(cfull(i), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), g.tmem_cols);
    }
    {   // conv weights -> smem, per (tap, channel half) a K-major 128B-swizzled [32][64] block
        const bf16 *w = static_cast<const bf16 *>(a.w);          // [32][3][128]
        for (int idx = threadIdx.x; idx < kCout * kTaps * kPlanes; idx += kThreads) {
            const int c = idx % kPlanes, t = (idx / kPlanes) % kTaps, n = idx / (kPlanes * kTaps);
            const int h = c >> 3, cc = c & 7;
            sts16(s_w + (uint32_t)((t * 2 + h) * 4096 + n * 128 + ((cc ^ (n & 7)) << 4)), ldg16(w + ((long long)n * kTaps + t) * kCin + c * 8));
        }
        // rows behind the last segment band are never written by TMA: zero them once (both buffers, both halves)
        const int rows = g.G * g.P, slack = g.px - rows;
        for (int idx = threadIdx.x; idx < 4 * slack * 8; idx += kThreads) {
            const int c = idx & 7, p = (idx >> 3) % slack, hb = (idx >> 3) / slack;
            sts16(s_slab0 + (uint32_t)hb * g.half_bytes + (uint32_t)(rows + p) * 128u + (uint32_t)c * 16u, make_uint4(0u, 0u, 0u, 0u));
        }
        // window indicator, K-major no-swizzle [k8 group][n = 16][8 rows]: 1 where slab row k is a frame of
        // (segment, window) n
        for (int idx = threadIdx.x; idx < g.k16 * 2 * kSumCols; idx += kThreads) {
            const int n = idx % kSumCols, k8 = idx / kSumCols;
            uint32_t v[4] = {0u, 0u, 0u, 0u};
            if (n < g.G * g.nwin) {
                const int gs = n / g.nwin, w = n - gs * g.nwin;
                const int r0 = gs * g.P + g.d + w * g.seg_len, r1 = gs * g.P + g.d + min(g.T, (w + 1) * g.seg_len);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int r = k8 * 8 + e;
                    if (r >= r0 && r < r1) v[e >> 1] |= (e & 1) ? 0x3F800000u : 0x00003F80u;       // bf16 1.0
                }
            }
            sts16(s_ind + (uint32_t)idx * 16u, make_uint4(v[0], v[1], v[2], v[3]));
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // everything above read only parameters; the activations come from the previous kernel

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (elect_one()) {
            const uint32_t band_bytes = (uint32_t)g.P * 128u;
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b0 = item * g.G;
                const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
                mbar_wait(sempty(buf), ph ^ 1u);
                CAM_TS(0);
                mbar_arrive_expect_tx(sfull(buf), 2u * (uint32_t)g.G * band_bytes);
                const uint32_t sb = s_slab0 + buf * g.slab_bytes;
                for (int gs = 0; gs < g.G; ++gs) {
                    tmap_load_3d(sb + (uint32_t)gs * band_bytes, &xmap, a.in_choff, -g.d, b0 + gs, sfull(buf));
                    tmap_load_3d(sb + g.half_bytes + (uint32_t)gs * band_bytes, &xmap, a.in_choff + 64, -g.d, b0 + gs, sfull(buf));
                }
                CAM_TS(1);
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        const uint32_t hi_sw = desc_hi_sw128(1024u);         // conv A and B: K-major SW128, 8-row groups 1024 B apart
        const uint32_t hi_ind = desc_hi(128u);               // indicator: K-major no swizzle
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(aempty(buf), ph ^ 1u);
            mbar_wait(sfull(buf), ph);
            tc_fence_after();
            CAM_TS(3);
            if (elect_one()) {
                const uint32_t sb = s_slab0 + buf * g.slab_bytes;
                const uint32_t tb = tmem_base + buf * acc_cols;
                {   // column sums first: the gate warps start while the conv MMAs run.  A = the slab read MN-major
                    // (64-channel groups one half apart, 8-row K groups 1024 B apart).  Even and odd K steps go to
                    // two accumulators; the gate threads add them.
                    uint32_t lo_a = desc_lo(sb, g.half_bytes), lo_b = desc_lo(s_ind, 256u);
                    const uint32_t dsum = tb + (uint32_t)g.n_tiles * 32u;
                    umma_bf16(dsum, desc64(lo_a, hi_sw), desc64(lo_b, hi_ind), kIdescSum, 0u);
                    if (g.k16 > 1) umma_bf16(dsum + kSumCols, desc64(lo_a + (2048u >> 4), hi_sw), desc64(lo_b + (512u >> 4), hi_ind), kIdescSum, 0u);
                    for (int k = 2; k < g.k16; ++k) {
                        lo_a += 2048u >> 4;                  // 16 rows of 128 B
                        lo_b += 512u >> 4;                   // two k8 groups of 16 x 16 B
                        umma_bf16(dsum + (uint32_t)(k & 1) * kSumCols, desc64(lo_a + (2048u >> 4), hi_sw), desc64(lo_b + (512u >> 4), hi_ind), kIdescSum, 1u);
                    }
                    umma_commit(cfull(buf));
                }
                // conv: per 128-row tile, 3 taps x 2 halves x 4 K steps, fully unrolled
                for (int t = 0; t < g.n_tiles; ++t) {
                    const uint32_t dcol = tb + (uint32_t)t * 32u;
#pragma unroll
                    for (int k = 0; k < kTaps; ++k) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t lo_a = desc_lo(sb + (uint32_t)h * g.half_bytes + (uint32_t)(t * 128 + k * g.d) * 128u, 16u);
                            const uint32_t lo_b = desc_lo(s_w + (uint32_t)((k * 2 + h) * 4096), 16u);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (k == 0 && h == 0 && j == 0) umma_bf16(dcol, desc64(lo_a, hi_sw), desc64(lo_b, hi_sw), kIdescN32, 0u);
                                else umma_acc(dcol, desc64(lo_a + (uint32_t)j * 2u, hi_sw), desc64(lo_b + (uint32_t)j * 2u, hi_sw), kIdescN32);
                            }
                        }
                    }
                }
                umma_commit(sempty(buf));
                umma_commit(afull(buf));
            }
            __syncwarp();
            CAM_TS(4);
        }
    } else if (warp < 6) {
        // =========================== context gate (warps 2-5, 128 threads) ===========================
        // thread = one input channel: TMEM lane (warp & 3) * 32 + lane of the column-sum accumulator
        const int gt = threadIdx.x - 64;                    // 0..127, the MLP work index
        const int q = warp & 3, ch = q * 32 + lane;
        // this thread's slice of the gate MLP, in registers for the life of the CTA (see gate_mlp)
        float2 w1r[32];
        float w2r[16];
        float2 b1r;
        float b2r;
        {
            const int ksl = lane >> 3, wq = gt >> 5;
            const int j = 16 * wq + 2 * (lane & 7), o = 8 * wq + (lane & 7);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int chn = 4 * i + ksl;
                w1r[i] = make_float2(j < g.hidden ? __ldg(w1t + chn * g.hidden + j) : 0.f,
                                     j + 1 < g.hidden ? __ldg(w1t + chn * g.hidden + j + 1) : 0.f);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int hu = 4 * i + ksl;
                w2r[i] = hu < g.hidden ? __ldg(w2t + hu * kCout + o) : 0.f;
            }
            b1r = make_float2(j < g.hidden ? __ldg(b1 + j) : 0.f, j + 1 < g.hidden ? __ldg(b1 + j + 1) : 0.f);
            b2r = __ldg(b2 + o);
        }
        const int nc = g.G * g.nwin;                        // (segment, window) combos of a full item
        const float inv_T = 1.f / (float)g.T, inv_seg = 1.f / (float)g.seg_len;
        const float inv_last = 1.f / (float)(g.T - (g.nwin - 1) * g.seg_len);
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            float *gateb = gate + (size_t)buf * g.G * g.nwin * kCout;
            mbar_wait(cfull(buf), ph);
            tc_fence_after();
            if (gt == 0) CAM_TS(5);
            {   // ---- this thread's channel: window sums from TMEM -> contexts
                uint32_t rr[16], ro[16];
                const uint32_t ts = tmem_base + buf * acc_cols + (uint32_t)g.n_tiles * 32u + ((uint32_t)(q * 32) << 16);
                tmem_ld16(ts, rr);
                tmem_ld16(ts + kSumCols, ro);
                tmem_ld_wait();
                tc_fence_before();
                if (gt == 0) CAM_TS(2);
                float sums[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) sums[e] = __uint_as_float(rr[e]) + (g.k16 > 1 ? __uint_as_float(ro[e]) : 0.f);
                switch (g.nwin) {
                    case 1: ctx_from_sums<1>(sums, win, g.G, ch, inv_T, inv_seg, inv_last); break;
                    case 2: ctx_from_sums<2>(sums, win, g.G, ch, inv_T, inv_seg, inv_last); break;
                    case 3: ctx_from_sums<3>(sums, win, g.G, ch, inv_T, inv_seg, inv_last); break;
                    default: ctx_from_sums<4>(sums, win, g.G, ch, inv_T, inv_seg, inv_last); break;
                }
            }
            if (gt == 0) CAM_TS(6);
            mbar_wait(gempty(buf), ph ^ 1u);                // the epilogue of item i-2 has read this gate buffer
            gate_bar_sync();
            if (gt == 0) CAM_TS(7);
            switch (nc) {
                case 1: gate_mlp<1>(win, hid, gateb, w1r, w2r, b1r, b2r, g.hidden, gt); break;
                case 2: gate_mlp<2>(win, hid, gateb, w1r, w2r, b1r, b2r, g.hidden, gt); break;
                case 3: gate_mlp<3>(win, hid, gateb, w1r, w2r, b1r, b2r, g.hidden, gt); break;
                case 4: gate_mlp<4>(win, hid, gateb, w1r, w2r, b1r, b2r, g.hidden, gt); break;
                case 5: gate_mlp<5>(win, hid, gateb, w1r, w2r, b1r, b2r, g.hidden, gt); break;
                case 6: gate_mlp<6>(win, hid, gateb, w1r, w2r, b1r, b2r, g.hidden, gt); break;
                case 7: gate_mlp<7>(win, hid, gateb, w1r, w2r, b1r, b2r, g.hidden, gt); break;
                default: gate_mlp<8>(win, hid, gateb, w1r, w2r, b1r, b2r, g.hidden, gt); break;
            }
            mbar_arrive(gfull(buf));
            if (gt == 0) CAM_TS(8);
        }
    } else {
        // =========================== epilogue (warps 6-9, 128 threads) ===========================
        const int q = warp & 3;                              // warps 6..9 -> TMEM lane quarters 2,3,0,1
        bf16 *y = static_cast<bf16 *>(a.y);
        const bool wide_st = (reinterpret_cast<uintptr_t>(a.y) & 31) == 0 && a.out_ld % 16 == 0 && a.out_choff % 16 == 0;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b0 = item * g.G;
            const int g_valid = min(g.G, a.B - b0);
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            const float *gateb = gate + (size_t)buf * g.G * g.nwin * kCout;
            mbar_wait(gfull(buf), ph);
            mbar_wait(afull(buf), ph);
            tc_fence_after();
            if (warp == 6) CAM_TS(9);
            for (int t = 0; t < g.n_tiles; ++t) {
                const int r = t * 128 + q * 32 + lane;
                const int gs = (int)__umulhi((unsigned)r, g.p_magic);
                const int u = r - gs * g.P;
                const bool ok = gs < g_valid && u < g.T;
                const uint32_t taddr = tmem_base + buf * acc_cols + (uint32_t)t * 32u + ((uint32_t)(q * 32) << 16);
                uint32_t rr[32];
                {
                    uint32_t (&r0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&rr[0]);
                    uint32_t (&r1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&rr[16]);
                    tmem_ld16(taddr, r0);
                    tmem_ld16(taddr + 16, r1);
                    tmem_ld_wait();
                }
                if (ok) {
                    const float *gr = &gateb[(gs * g.nwin + u / g.seg_len) * kCout];
                    float v[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(rr[e]) * gr[e];
                    bf16 *yp = y + ((long long)(b0 + gs) * g.T + u) * a.out_ld + a.out_choff;
                    if (wide_st) {      // two full 32-byte sectors per lane instead of four half-sector stores
#pragma unroll
                        for (int e = 0; e < 32; e += 16)
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yp + e), "r"(pack2(v[e], v[e + 1])),
                                         "r"(pack2(v[e + 2], v[e + 3])), "r"(pack2(v[e + 4], v[e + 5])), "r"(pack2(v[e + 6], v[e + 7])),
                                         "r"(pack2(v[e + 8], v[e + 9])), "r"(pack2(v[e + 10], v[e + 11])), "r"(pack2(v[e + 12], v[e + 13])),
                                         "r"(pack2(v[e + 14], v[e + 15])) : "memory");
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; e += 8)
                            *reinterpret_cast<uint4 *>(yp + e) =
                                make_uint4(pack2(v[e], v[e + 1]), pack2(v[e + 2], v[e + 3]), pack2(v[e + 4], v[e + 5]), pack2(v[e + 6], v[e + 7]));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(aempty(buf));
            mbar_arrive(gempty(buf));
            if (warp == 6) CAM_TS(10);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, g.tmem_cols);
}

bool geometry(const ConvArgs &a, int hidden, int seg_len, CamGeom &g) {
    g.T = a.W; g.d = a.dw; g.seg_len = seg_len; g.hidden = hidden;
    g.nwin = (g.T + seg_len - 1) / seg_len;
    g.P = (g.T + 2 * g.d + 7) & ~7;             // band pitch: a multiple of the 8-row swizzle atom
    if (g.P > 256) return false;                // one TMA box per segment and channel half
    g.G = std::min(kMaxSeg, 248 / g.P);
    if (g.G < 1 || g.nwin > kMaxWin || hidden > kMaxHidden || hidden < 4 || hidden % 4) return false;
    while (g.G * g.nwin > 8) --g.G;             // the gate MLP keeps at most 8 (segment, window) combos in registers
    g.n_tiles = (g.G * g.P + 127) / 128;
    // rows the shifted views can touch: n_tiles*128 + 2d, rounded to whole swizzle atoms
    g.px = (std::max(g.G * g.P, g.n_tiles * 128) + 2 * g.d + 7) & ~7;
    g.p_magic = (unsigned)(((1ull << 32) + g.P - 1) / g.P);
    g.half_bytes = (uint32_t)g.px * 128u;
    g.slab_bytes = 2u * g.half_bytes;
    g.k16 = (g.G * g.P + 15) / 16;
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off += (bytes + 1023u) & ~1023u; return o; };
    g.off_w = take(kTaps * 2 * 4096);
    g.off_slab = take(2 * g.slab_bytes);
    g.off_ind = take(g.k16 * 2 * kSumCols * 16);
    g.off_mlp = 0;
    g.off_part = 0;
    g.off_win = off; off += g.G * g.nwin * kCin * 4;
    g.off_hid = off; off += 8 * kMaxHidden * 4;
    g.off_gate = off; off += 2 * g.G * g.nwin * kCout * 4;
    off = (off + 127u) & ~127u;
    g.off_bar = off; off += 256;
    g.smem_bytes = (int)off;
    const int cols = 2 * (g.n_tiles * 32 + 2 * kSumCols);
    g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    return g.smem_bytes <= 227 * 1024 && cols <= 512;
}

// {64 channels, P frames, 1 segment} boxes over the [B][T][ld] activation buffer, 128B swizzle
int input_map(const ConvArgs &a, const CamGeom &g, CUtensorMap *out) {
    typedef std::tuple<const void *, int, int, int, int> Key;
    static std::mutex mu;
    static std::map<Key, CUtensorMap> cache;
    const Key key(a.x, a.in_ld, g.T, a.B, g.P);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return SPK_OK;
    }
    const uint64_t dims[3] = {(uint64_t)a.in_ld, (uint64_t)g.T, (uint64_t)a.B};
    const uint64_t strides[2] = {(uint64_t)a.in_ld * 2, (uint64_t)g.T * a.in_ld * 2};
    const uint32_t box[3] = {64u, (uint32_t)g.P, 1u};
    const int rc = tmap_encode_bf16(a.x, 3, dims, strides, box, 128, out);
    if (rc != SPK_OK) return rc;
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return SPK_OK;
}

}  // namespace

bool cam_local_supported(const ConvArgs &a, int in_dtype, int out_dtype, int hidden, int seg_len) {
    if (in_dtype != SPK_DT_BF16 || out_dtype != SPK_DT_BF16) return false;
    if (a.Cin != kCin || a.Cout != kCout || a.KH != 1 || a.KW != kTaps || a.H != 1 || a.Ho != 1) return false;
    if (a.sw != 1 || a.sh != 1 || a.pw != a.dw || a.dw < 1 || a.Wo != a.W) return false;
    if (a.pro_scale != nullptr || a.epi_scale != nullptr || a.res != nullptr || a.act != SPK_ACT_NONE) return false;
    if (a.post_scale != nullptr || a.pad_reflect) return false;
    if (a.in_ld % 8 || a.in_choff % 8 || a.out_ld % 8 || a.out_choff % 8) return false;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.y) & 15)) return false;
    if (seg_len < 1) return false;
    CamGeom g;
    return geometry(a, hidden, seg_len, g);
}

int launch_cam_local(const ConvArgs &a, const float *w1t, const float *b1, const float *w2t, const float *b2, int hidden,
                     int seg_len, cudaStream_t s) {
    CamGeom g;
    if (!geometry(a, hidden, seg_len, g)) {
        set_error("cam_local: geometry does not fit");
        return SPK_ERR_UNSUPPORTED;
    }
    if (a.B == 0) return SPK_OK;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(cam_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(cam_local) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    CUtensorMap xmap;
    const int rc = input_map(a, g, &xmap);
    if (rc != SPK_OK) return rc;
    const int items = (a.B + g.G - 1) / g.G;
    const int grid = std::min(items, sm_count());
    static const int dbg = getenv("SPK_CAM_DBG") ? atoi(getenv("SPK_CAM_DBG")) : 0;
    const cudaError_t le = launch_pdl(cam_local_kernel, dim3(grid), dim3(kThreads), (size_t)g.smem_bytes, s, a, g, w1t, b1, w2t, b2, items, dbg, xmap);
    if (le != cudaSuccess) {
        set_error("cam_local_kernel launch failed: %s", cudaGetErrorString(le));
        return SPK_ERR_CUDA;
    }
    return check_launch("cam_local_kernel");
}

}  // namespace spk

// debug aid (not part of the ABI): per-item role timestamps of CTA 0 of the last launch with SPK_CAM_DBG & 64
extern "C" int spk_debug_cam_timeline(long long *dst) {
    return cudaMemcpyFromSymbol(dst, spk::g_cam_ts, sizeof(long long) * 16 * 12) == cudaSuccess ? 0 : -1;
}
