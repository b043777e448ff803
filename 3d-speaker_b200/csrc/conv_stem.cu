// Stem + first FCM block entry of CAM++ in one kernel (speakerlab/models/campplus/DTDNN.py:39-48, layers.py:221-253):
//
//     s  = relu(bn1(conv3x3(feats)))            1 -> 32 channels on the [F, T] image      (FCM.conv1 / bn1)
//     y1 = relu(bn(conv3x3 stride (2,1) (s)))   first conv of layer1[0]                   (BasicResBlock.conv1 / bn1)
//     y2 = bn(conv1x1 stride (2,1) (s))         its shortcut                              (BasicResBlock.shortcut)
//
// The stem output is the largest tensor of the network (757 KB per 1.5 s segment in bf16), written once and read
// twice.  Here it never reaches HBM: for a band of R output rows the 2R + 1 stem rows it needs are computed by eight
// CUDA-core warps straight into the staged slab of the stride-2 slab kernel (conv_slab3.cu: even / odd input rows in
// two 64B-swizzled sub-slabs, zero padding = pixels nobody writes), the 3x3 conv and the shortcut run as tcgen05 MMAs
// on pixel-shifted views of that slab into two TMEM accumulators, and both outputs leave through swizzled staging
// buffers and TMA box stores.  Input traffic is the 47 KB of features per segment.
//
//   warps 0-7   stem: feature tile -> shared memory (transposed, double buffered, the loads of item i+2 in registers);
//               the 1 -> 32 conv itself runs on the tensor core too: a thread builds the im2col row of ITS pixel as split
//               bf16 - [x_hi (9) | x_lo (9) | x_hi (9) | 1 | 1] against [w_hi | w_hi | w_lo | shift_hi | shift_lo] (BN scale
//               folded into w) is the fp32 product to ~2^-17, K = 32 - 256 pixels per step, one elected thread issues the
//               four MMAs of the step, and every thread then reads its pixel's 32 channels back from TMEM, applies ReLU
//               and writes the bf16 pixel into the slab.  (On CUDA cores the same conv is 36 FFMA2 of ~105 instructions per
//               pixel and 8 channels: ~6,000 cycles per item with the eight warps one CTA per SM can spare - more than the
//               three separate kernels took.)
//   warp 8      MMA issuer (+ TMEM allocation): 18 MMAs per 128-pixel tile for the conv, 2 for the shortcut
//   warps 9-12  epilogue: BN (+ReLU) of both accumulators -> staging buffers
//   warp 13     TMA stores
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

using namespace tc;
using bf16 = __nv_bfloat16;
constexpr int kC = 32;
constexpr int kStemWarps = 8;
constexpr int kStemThreads = kStemWarps * 32, kEpiThreads = 128;
constexpr int kThreads = kStemThreads + 32 + kEpiThreads + 32;      // 448
constexpr int kStemTries = 1;                                       // polls for a pending stem step in front of every conv tile
constexpr int kMaxLd = 7;                                           // feature loads per stem thread and item
constexpr uint32_t kIdesc = idesc_bf16(32);

struct StemGeom {
    int Wp, R, n_tiles, n_bands, rows_e, rows_o, px_e, px_o, nbins, Tp2;
    uint32_t off_w1, off_ws, off_sb, off_ss, off_feat, off_sa, off_stg, stg_bytes, off_slab, slab_bytes, off_bar;
    int smem_bytes, tmem_cols;
    unsigned t_magic;        // ceil(2^32 / T)
};

// debug aid: per-item role timestamps of CTA 0 (spk_debug_stem_enable(1)), read back by spk_debug_stem_timeline
__device__ long long g_stem_ts[64 * 8];
__device__ int g_stem_dbg;
#define STEM_TS(idx, slot) do { if (dbg && (idx) < 64) g_stem_ts[(idx) * 8 + (slot)] = clock64(); } while (0)

// (lo, hi) -> bf16x2 with ReLU in the conversion
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// v -> bf16(v) | bf16(v - bf16(v)) << 16: the split-bf16 form of one feature value
__device__ __forceinline__ uint32_t split_bf16(float v) {
    const bf16 h = __float2bfloat16_rn(v);
    const bf16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    return (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ void stem_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kStemThreads) : "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
stem_block_kernel(const StemBlockArgs a, const StemGeom g, long long n_items, const __grid_constant__ CUtensorMap y1map,
                  const __grid_constant__ CUtensorMap y2map) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t s0 = smem_u32(smem);
    const uint32_t s_w1 = s0 + g.off_w1, s_ws = s0 + g.off_ws, s_ss = s0 + g.off_ss, s_slab0 = s0 + g.off_slab, s_bar = s0 + g.off_bar;
    const uint32_t s_stg = s0 + g.off_stg, s_sa = s0 + g.off_sa, s_sb = s0 + g.off_sb;
    uint32_t *feat = reinterpret_cast<uint32_t *>(smem + g.off_feat);    // 2 x [nbins][Tp2] split-bf16 values (hi | lo << 16), columns 0 and T + 1 stay zero
    auto sfull = [&](uint32_t i) { return s_bar + 8u * i; };
    auto sempty = [&](uint32_t i) { return s_bar + 8u * (2 + i); };
    auto afull = [&](uint32_t i) { return s_bar + 8u * (4 + i); };
    auto aempty = [&](uint32_t i) { return s_bar + 8u * (6 + i); };
    auto gfull = [&](uint32_t o) { return s_bar + 8u * (8 + o); };        // staging buffer of output o (0 conv, 1 shortcut) is complete
    auto gfree = [&](uint32_t o) { return s_bar + 8u * (14 + o); };       // its TMA store has read it
    auto smma = [&](uint32_t i) { return s_bar + 8u * (10 + i); };        // stem GEMM: the MMAs of a step have retired
    auto aready = [&](uint32_t i) { return s_bar + 8u * (12 + i); };      //            the A rows of a step are staged
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + g.off_bar + 136);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool dbg = g_stem_dbg != 0 && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (warp == 0 || warp == kStemWarps || warp == kStemWarps + 1 || warp == kStemWarps + 5);
    int di = 0;
    const uint32_t sub_o = (uint32_t)g.px_e * 64u;          // byte offset of the odd sub-slab inside a slab buffer
    const uint32_t acc_cols = (uint32_t)g.n_tiles * 32u;    // one accumulator set; a buffer holds two (conv, shortcut)

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(sfull(i), kStemThreads / 32);
            mbar_init(sempty(i), 1);
            mbar_init(afull(i), 1);
            mbar_init(aempty(i), kEpiThreads);
        }
        for (uint32_t o = 0; o < 2; ++o) {
            mbar_init(gfull(o), kEpiThreads);
            mbar_init(gfree(o), 1);
        }
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(smma(i), 1);
            mbar_init(aready(i), kStemWarps);
        }
        tmap_prefetch(&y1map);
        tmap_prefetch(&y2map);
        fence_barrier_init();
    }
    if (warp == kStemWarps) {
        __syncwarp();
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), g.tmem_cols);
    }
    {   // conv weights -> smem: per tap a K-major 64B-swizzled [32 cout][32 cin] block (chunk c of row n at c ^ ((n >> 1) & 3));
        // BN vectors; the whole slab area and the feature tile start as zeros (pad columns / slack pixels are never written)
        for (int idx = threadIdx.x; idx < kC * 9 * 4; idx += kThreads) {
            const int c = idx & 3, t = (idx >> 2) % 9, n = idx / 36;
            sts16(s_w1 + (uint32_t)(t * 2048 + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)), ldg16(a.w1 + ((long long)n * 9 + t) * kC + c * 8));
        }
        for (int idx = threadIdx.x; idx < kC * 4; idx += kThreads) {
            const int c = idx & 3, n = idx >> 2;
            sts16(s_ws + (uint32_t)(n * 64 + ((c ^ ((n >> 1) & 3)) << 4)), ldg16(a.ws + (long long)n * kC + c * 8));
        }
        // stem conv as a K = 32 GEMM operand: row n = [w_hi (9) | w_hi (9) | w_lo (9) | shift_hi | shift_lo | 0 0 0], w = weight x BN scale
        for (int idx = threadIdx.x; idx < kC * 32; idx += kThreads) {
            const int n = idx >> 5, k = idx & 31;
            float v = 0.f;
            bool lo = false;
            if (k < 27) {
                v = __ldg(a.w0 + n * 9 + (k % 9)) * __ldg(a.s0 + n);
                lo = k >= 18;
            } else if (k < 29) {
                v = __ldg(a.b0 + n);
                lo = k == 28;
            }
            const bf16 hi = __float2bfloat16_rn(v);
            const bf16 out = lo ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
            *reinterpret_cast<bf16 *>(smem + g.off_sb + n * 64 + (((k >> 3) ^ ((n >> 1) & 3)) << 4) + (k & 7) * 2) = out;
        }
        float *ss = reinterpret_cast<float *>(smem + g.off_ss);
        for (int n = threadIdx.x; n < kC; n += kThreads) {
            ss[n] = __ldg(a.s1 + n); ss[32 + n] = __ldg(a.b1 + n);
            ss[64 + n] = __ldg(a.ss + n); ss[96 + n] = __ldg(a.bs + n);
        }
        for (uint32_t off = (uint32_t)threadIdx.x * 16u; off < 2u * g.slab_bytes; off += kThreads * 16u)
            sts16(s_slab0 + off, make_uint4(0u, 0u, 0u, 0u));
        for (int i = threadIdx.x; i < 2 * g.nbins * g.Tp2; i += kThreads) feat[i] = 0u;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // no pdl_wait before the feature loads: feats come from the fbank kernel, which is not a PDL-overlapped predecessor
    pdl_wait();

    if (warp < kStemWarps) {
        // =========================== stem ===========================
        const int tid = threadIdx.x;
        const int tile = tid >> 7, r = tid & 127;          // this thread's A row / TMEM lane in a 256-pixel step
        const uint32_t arow = s_sa + (uint32_t)tile * 8192u + (uint32_t)r * 64u, ax = ((uint32_t)r >> 1) & 3u;
        const uint32_t t_stem = tmem_base + 4u * acc_cols + (uint32_t)tile * 32u + ((uint32_t)((warp & 3) * 32) << 16);
        const int n_feat = g.nbins * a.T;
        const int n_px = (g.rows_e + g.rows_o) * a.T;
        const int n_steps = (n_px + kStemThreads - 1) / kStemThreads;
        const int tile_f = g.nbins * g.Tp2;
        // this thread's feature elements: (frame, bin) -> global offset, shared-memory offset, bin (the same for every item)
        int goff[kMaxLd], soff[kMaxLd];
#pragma unroll
        for (int u = 0; u < kMaxLd; ++u) {
            const int e = tid + u * kStemThreads;
            const int t = e / g.nbins, i = e - t * g.nbins;
            goff[u] = e < n_feat ? t * a.F + i : -1;
            soff[u] = (i * g.Tp2 + t + 1) | (i << 20);
        }
        float pre[kMaxLd];
        auto prefetch = [&](long long item) {
            const int b = (int)(item / g.n_bands);
            const int band = (int)(item - (long long)b * g.n_bands);
            const int bin0 = 2 * band * g.R - 2;
            const float *fe = a.feats + (size_t)b * a.T * a.F + bin0;
#pragma unroll
            for (int u = 0; u < kMaxLd; ++u) {
                const int f = bin0 + (soff[u] >> 20);
                pre[u] = (goff[u] >= 0 && f >= 0 && f < a.F) ? __ldg(fe + goff[u]) : 0.f;
            }
        };
        auto stage = [&](uint32_t *dst) {        // every feature value is split once here, not nine times in the im2col rows
#pragma unroll
            for (int u = 0; u < kMaxLd; ++u)
                if (goff[u] >= 0) dst[soff[u] & 0xFFFFF] = split_bf16(pre[u]);
        };
        // ---- one 256-pixel step of the stem GEMM, in two halves that are software-pipelined ACROSS steps (and items):
        // build(step n + 1) runs before finish(step n), so the MMA round trip of a step hides behind the next step's rows.
        //   build : im2col row of this thread's pixel (split bf16) -> A buffer (gs & 1), arrive
        //   finish: wait for the step's MMAs, read the pixel's 32 channels from TMEM, ReLU, bf16 -> slab
        struct Step { long long item; int s; uint32_t gs; };
        auto build = [&](const Step &st, const uint32_t *ft) {
            const int q = st.s * kStemThreads + tid;
            if (q < n_px) {
                const int j = (int)__umulhi((unsigned)q, g.t_magic);         // q / T  (exact for q < 2^16)
                const int t = q - j * a.T;
                const uint32_t *f0 = ft + j * g.Tp2 + t, *f1 = f0 + g.Tp2, *f2 = f1 + g.Tp2;
                const uint32_t in[9] = {f0[0], f0[1], f0[2], f1[0], f1[1], f1[2], f2[0], f2[1], f2[2]};     // hi | lo << 16
                // row = [hi0..8 | lo0..8 | hi0..8 | 1 | 1 | 0 0 0] as sixteen bf16 pairs, assembled with byte permutes
                constexpr uint32_t LL = 0x5410, HH = 0x7632, LH = 0x7610;       // (a.lo, b.lo), (a.hi, b.hi), (a.lo, b.hi)
                constexpr uint32_t one = 0x3F80;
                const uint32_t h01 = prmt(in[0], in[1], LL), h23 = prmt(in[2], in[3], LL), h45 = prmt(in[4], in[5], LL), h67 = prmt(in[6], in[7], LL);
                const uint32_t ar = arow + (st.gs & 1u) * 16384u;
                sts16(ar + ((0u ^ ax) << 4), make_uint4(h01, h23, h45, h67));
                sts16(ar + ((1u ^ ax) << 4), make_uint4(prmt(in[8], in[0], LH), prmt(in[1], in[2], HH), prmt(in[3], in[4], HH), prmt(in[5], in[6], HH)));
                sts16(ar + ((2u ^ ax) << 4), make_uint4(prmt(in[7], in[8], HH), h01, h23, h45));
                sts16(ar + ((3u ^ ax) << 4), make_uint4(h67, (in[8] & 0xFFFFu) | (one << 16), one, 0u));
            }
            tc_fence_before();              // this thread's TMEM reads of step gs - 2 (same accumulator) are done
            fence_proxy_async();            // A rows (generic proxy) -> tcgen05.mma (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(aready(st.gs & 1u));      // the MMA warp issues the step's four MMAs between conv tiles
        };
        uint32_t buf = 0, ph = 0;           // slab buffer / phase of the item being FINISHED
        auto finish = [&](const Step &st) {
            const int band = (int)(st.item % g.n_bands);
            const int xr0 = 2 * band * g.R - 1;              // stem row of slab row j = 0
            if (st.s == 0) {
                STEM_TS(di, 0);
                mbar_wait(sempty(buf), ph ^ 1u);
                STEM_TS(di, 1);
            }
            const uint32_t sb = s_slab0 + buf * g.slab_bytes;
            mbar_wait(smma(st.gs & 1u), (st.gs >> 1) & 1u);
            tc_fence_after();
            uint32_t v[32];
            {
                uint32_t (&v0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&v[0]);
                uint32_t (&v1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&v[16]);
                const uint32_t ta = t_stem + (st.gs & 1u) * 64u;
                tmem_ld16(ta, v0);
                tmem_ld16(ta + 16, v1);
                tmem_ld_wait();
            }
            const int q = st.s * kStemThreads + tid;
            if (q < n_px) {
                const int j = (int)__umulhi((unsigned)q, g.t_magic);
                const int t = q - j * a.T;
                const int xr = xr0 + j;
                const uint32_t p = (uint32_t)((j >> 1) * g.Wp + t + 1);
                const uint32_t prow = sb + ((j & 1) ? sub_o : 0u) + p * 64u, x = (p >> 1) & 3u;
                const bool valid = xr >= 0 && xr < a.F;          // else: conv zero padding above / below the image
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (valid) {
                        auto f = [&](int i) { return __uint_as_float(v[e * 8 + i]); };
                        o = make_uint4(pack2_relu(f(0), f(1)), pack2_relu(f(2), f(3)), pack2_relu(f(4), f(5)), pack2_relu(f(6), f(7)));
                    }
                    sts16(prow + (((uint32_t)e ^ x) << 4), o);
                }
            }
            if (st.s == n_steps - 1) {
                fence_proxy_async();            // slab writes (generic proxy) -> tcgen05.mma reads (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(sfull(buf));
                STEM_TS(di, 2);
                ++di;
                if (++buf == 2u) { buf = 0; ph ^= 1u; }
            }
        };
        const long long step = gridDim.x;
        uint32_t fb = 0, gs = 0;
        Step prev{0, 0, 0};
        bool have_prev = false;
        if ((long long)blockIdx.x < n_items) {
            prefetch(blockIdx.x);
            stage(feat);
            if (blockIdx.x + step < n_items) prefetch(blockIdx.x + step);
        }
        for (long long item = blockIdx.x; item < n_items; item += step) {
            // entering an item: its tile was staged one item ago; after the barrier nobody reads the other tile any more,
            // so the next item's tile (in registers) goes there and the loads of the item after that are issued
            stem_bar_sync();
            const uint32_t *ft = feat + fb * tile_f;
            if (item + step < n_items) {
                stage(feat + (fb ^ 1u) * tile_f);
                if (item + 2 * step < n_items) prefetch(item + 2 * step);
            }
            for (int sidx = 0; sidx < n_steps; ++sidx, ++gs) {
                const Step cur{item, sidx, gs};
                build(cur, ft);
                if (have_prev) finish(prev);
                prev = cur;
                have_prev = true;
            }
            fb ^= 1u;
        }
        if (have_prev) finish(prev);
    } else if (warp == kStemWarps) {
        // =========================== MMA issuer ===========================
        const uint32_t hi = desc_hi(512u, kLayoutSw64);         // 8-pixel groups are 512 B apart
        const uint32_t wp4 = (uint32_t)g.Wp * 4u;               // one slab row, in 16-byte descriptor units
        const uint32_t d_stem = tmem_base + 4u * acc_cols;
        // The stem GEMM shares the tensor pipe with the conv: its four MMAs per 256-pixel step are issued HERE, between
        // the conv tiles and while waiting, so that a step never queues behind a whole item of conv MMAs (2,900 cycles).
        uint32_t sgs = 0;                    // stem steps served so far
        auto service = [&]() -> bool {
            const uint32_t ab = sgs & 1u;
            const bool ready = __shfl_sync(0xffffffffu, mbar_try_wait(aready(ab), (sgs >> 1) & 1u) ? 1 : 0, 0) != 0;
            if (!ready) return false;
            ++sgs;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t lo_b = desc_lo(s_sb, 16u);
#pragma unroll
                for (uint32_t tl = 0; tl < 2; ++tl) {
                    const uint32_t lo_a = desc_lo(s_sa + ab * 16384u + tl * 8192u, 16u);
                    const uint32_t dd = d_stem + ab * 64u + tl * 32u;
                    umma_bf16(dd, desc64(lo_a, hi), desc64(lo_b, hi), kIdesc, 0u);
                    umma_bf16_acc(dd, desc64(lo_a + 2u, hi), desc64(lo_b + 2u, hi), kIdesc);
                }
                umma_commit(smma(ab));
            }
            __syncwarp();
            return true;
        };
        auto wait_serving = [&](uint32_t bar, uint32_t parity) {
            uint32_t spins = 0;
            while (__shfl_sync(0xffffffffu, mbar_try_wait(bar, parity) ? 1 : 0, 0) == 0) {
                service();
                if (++spins > kSpinLimit) __trap();
            }
        };
        uint32_t it = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            wait_serving(aempty(buf), ph ^ 1u);
            wait_serving(sfull(buf), ph);
            tc_fence_after();
            STEM_TS(di, 3);
            const uint32_t lo_e = desc_lo(s_slab0 + buf * g.slab_bytes, 16u), lo_o = lo_e + (sub_o >> 4);
            const uint32_t lo_w = desc_lo(s_w1, 16u), lo_s = desc_lo(s_ws, 16u);
            uint32_t d = tmem_base + buf * 2u * acc_cols;
            uint32_t tile = 0;                                // 128 pixels = 8192 B = 512 units
            for (int t = 0; t < g.n_tiles; ++t, d += 32u, tile += 512u) {
                // one stem step (of the item the stem warps are building) in front of every conv tile: the step's CUDA-core
                // half then runs under the tile's 800 tensor cycles.  Bounded wait: there is no step left at the tail.
                for (int tries = 0; tries < kStemTries && !service(); ++tries) {}
                if (elect_one()) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const uint32_t lo_a = (kh & 1) ? lo_o + tile + 4u * kw : lo_e + tile + (kh >> 1) * wp4 + 4u * kw;
                            const uint32_t lo_b = lo_w + (uint32_t)((kh * 3 + kw) * (2048 >> 4));
                            if (kh == 0 && kw == 0) umma_bf16(d, desc64(lo_a, hi), desc64(lo_b, hi), kIdesc, 0u);
                            else umma_bf16_acc(d, desc64(lo_a, hi), desc64(lo_b, hi), kIdesc);
                            umma_bf16_acc(d, desc64(lo_a + 2u, hi), desc64(lo_b + 2u, hi), kIdesc);       // channels 16-31
                        }
                    // shortcut: the centre tap's pixels (even image rows) times the 1x1 weights -> second accumulator
                    const uint32_t lo_c = lo_o + tile + 4u;
                    umma_bf16(d + acc_cols, desc64(lo_c, hi), desc64(lo_s, hi), kIdesc, 0u);
                    umma_bf16_acc(d + acc_cols, desc64(lo_c + 2u, hi), desc64(lo_s + 2u, hi), kIdesc);
                    if (t == g.n_tiles - 1) {
                        umma_commit(sempty(buf));      // slab reusable once these MMAs retire
                        umma_commit(afull(buf));
                    }
                }
                __syncwarp();
            }
            STEM_TS(di, 4);
            ++di;
        }
    } else if (warp < kStemWarps + 5) {
        // =========================== epilogue ===========================
        const int q = warp & 3;
        const uint32_t band_px = (uint32_t)(g.R * g.Wp);                 // the staging buffers hold exactly the band
        uint32_t it = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(afull(buf), ph);
            tc_fence_after();
            STEM_TS(di, 5);
            // output by output (0: conv, BN + ReLU; 1: shortcut, BN), so that the store of the conv band runs under the
            // shortcut's epilogue and the next item only has to wait for the store that is one full output pass old
#pragma unroll
            for (uint32_t o = 0; o < 2; ++o) {
                mbar_wait(gfree(o), (it & 1u) ^ 1u);          // the previous item's store of this output has read the buffer
                const uint32_t tbase = tmem_base + buf * 2u * acc_cols + o * acc_cols + ((uint32_t)(q * 32) << 16);
                const uint32_t sbase = s_stg + o * g.stg_bytes;
                for (int t0 = 0; t0 < g.n_tiles; t0 += 2) {       // two tiles per TMEM round trip
                    const bool two = t0 + 1 < g.n_tiles;
                    uint32_t r[2][32];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (u == 0 || two) {
                            uint32_t (&r0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[u][0]);
                            uint32_t (&r1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[u][16]);
                            tmem_ld16(tbase + (uint32_t)(t0 + u) * 32u, r0);
                            tmem_ld16(tbase + (uint32_t)(t0 + u) * 32u + 16, r1);
                        }
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (u == 1 && !two) break;
                        const uint32_t p = (uint32_t)((t0 + u) * 128 + q * 32 + lane);        // slab pixel == staging pixel
                        if (p >= band_px) continue;                                            // tile overhang
                        const uint32_t x = (p >> 1) & 3u, prow = sbase + p * 64u;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float v[8];
#pragma unroll
                            for (int h = 0; h < 8; h += 4) {
                                const uint4 s4 = lds16(s_ss + (uint32_t)(o * 256 + (e * 8 + h) * 4)), h4 = lds16(s_ss + (uint32_t)(o * 256 + 128 + (e * 8 + h) * 4));
                                v[h] = fmaf(__uint_as_float(r[u][e * 8 + h]), __uint_as_float(s4.x), __uint_as_float(h4.x));
                                v[h + 1] = fmaf(__uint_as_float(r[u][e * 8 + h + 1]), __uint_as_float(s4.y), __uint_as_float(h4.y));
                                v[h + 2] = fmaf(__uint_as_float(r[u][e * 8 + h + 2]), __uint_as_float(s4.z), __uint_as_float(h4.z));
                                v[h + 3] = fmaf(__uint_as_float(r[u][e * 8 + h + 3]), __uint_as_float(s4.w), __uint_as_float(h4.w));
                            }
                            const uint4 out = o == 0 ? make_uint4(pack2_relu(v[0], v[1]), pack2_relu(v[2], v[3]), pack2_relu(v[4], v[5]), pack2_relu(v[6], v[7]))
                                                     : make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                            sts16(prow + (((uint32_t)e ^ x) << 4), out);
                        }
                    }
                }
                if (o == 1) {
                    tc_fence_before();
                    mbar_arrive(aempty(buf));       // both accumulators of the buffer have been read
                }
                fence_proxy_async();            // staging writes (generic proxy) -> TMA store (async proxy)
                mbar_arrive(gfull(o));
            }
            STEM_TS(di, 6);
            ++di;
        }
    } else {
        // =========================== TMA store ===========================
        if (elect_one()) {
            uint32_t it = 0;
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b = (int)(item / g.n_bands);
                const int band = (int)(item - (long long)b * g.n_bands);
                // the conv band leaves while the epilogue works on the shortcut band
                mbar_wait(gfull(0), it & 1u);
                tmap_store_4d(&y1map, a.y1_choff, 0, band * g.R, b, s_stg);
                bulk_commit();
                mbar_wait(gfull(1), it & 1u);
                tmap_store_4d(&y2map, a.y2_choff, 0, band * g.R, b, s_stg + g.stg_bytes);
                bulk_commit();
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the first box has been read out
                mbar_arrive(gfree(0));
                bulk_wait_read0();
                mbar_arrive(gfree(1));
                STEM_TS(di, 7);
                ++di;
            }
            bulk_wait_all();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kStemWarps) tmem_dealloc(tmem_base, g.tmem_cols);
}

bool geometry(const StemBlockArgs &a, StemGeom &g) {
    g.Wp = (a.T + 2 + 7) & ~7;
    if (g.Wp > 256 || a.F % 2 != 0 || a.Ho != a.F / 2) return false;
    for (int R = std::min(4, a.Ho); R >= 1; --R) {
        g.R = R;
        g.n_tiles = (R * g.Wp + 127) / 128;
        g.rows_e = R + 1;
        g.rows_o = R;
        g.px_e = (std::max(g.rows_e * g.Wp, g.n_tiles * 128 + g.Wp + 2) + 7) & ~7;
        g.px_o = (std::max(g.rows_o * g.Wp, g.n_tiles * 128 + 2) + 7) & ~7;
        g.slab_bytes = (64u * (uint32_t)(g.px_e + g.px_o) + 1023u) & ~1023u;
        g.stg_bytes = ((uint32_t)(R * g.Wp) * 64u + 1023u) & ~1023u;       // exactly the band (the epilogue masks the tile overhang)
        g.nbins = 2 * R + 3;
        g.Tp2 = a.T + 2;
        g.off_w1 = 0;
        g.off_ws = 9 * 2048;
        g.off_sb = 10 * 2048;
        g.off_ss = 11 * 2048;
        g.off_feat = g.off_ss + 512;
        g.off_sa = (g.off_feat + 2u * (uint32_t)(g.nbins * g.Tp2 * 4) + 1023u) & ~1023u;
        g.off_stg = g.off_sa + 4u * 8192u;              // two A buffers of two 128-pixel tiles
        g.off_slab = g.off_stg + 2u * g.stg_bytes;
        g.off_bar = g.off_slab + 2u * g.slab_bytes;
        g.smem_bytes = (int)g.off_bar + 176;
        const int cols = g.n_tiles * 32 * 4 + 128;      // two accumulator sets, two buffers; 2 x 2 stem accumulators
        g.tmem_cols = cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
        if (cols <= 512 && g.smem_bytes <= 227 * 1024 && g.nbins * a.T <= kMaxLd * kStemThreads) {
            g.n_bands = (a.Ho + R - 1) / R;
            g.t_magic = (unsigned)(((1ull << 32) + a.T - 1) / a.T);
            return true;
        }
    }
    return false;
}

int out_map(const void *ptr, int ld, const StemBlockArgs &a, int wp, int rows, CUtensorMap *out) {
    typedef std::tuple<const void *, int, int, int, int, int, int> Key;
    static std::mutex mu;
    static std::map<Key, CUtensorMap> cache;
    const Key key(ptr, ld, a.T, a.Ho, a.B, wp, rows);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return SPK_OK;
    }
    const uint64_t dims[4] = {(uint64_t)ld, (uint64_t)a.T, (uint64_t)a.Ho, (uint64_t)a.B};
    const uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)a.T * ld * 2, (uint64_t)a.Ho * a.T * ld * 2};
    const uint32_t box[4] = {32u, (uint32_t)wp, (uint32_t)rows, 1u};
    const int rc = tmap_encode_bf16(ptr, 4, dims, strides, box, 64, out);
    if (rc != SPK_OK) return rc;
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return SPK_OK;
}

}  // namespace

bool stem_block_supported(const StemBlockArgs &a) {
    static const bool off = [] { const char *e = getenv("SPK_NO_STEM_FUSE"); return e && e[0] == '1'; }();
    if (off || a.T < 8 || a.F < 4) return false;
    if (a.y1_ld % 8 || a.y1_choff % 8 || a.y2_ld % 8 || a.y2_choff % 8) return false;
    if ((reinterpret_cast<uintptr_t>(a.y1) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.y2) & 15) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.w1) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.ws) & 15) != 0) return false;
    StemGeom g;
    return geometry(a, g);
}

int launch_stem_block(const StemBlockArgs &a, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    StemGeom g;
    if (!geometry(a, g)) {
        set_error("stem_block: geometry does not fit (T=%d, F=%d)", a.T, a.F);
        return SPK_ERR_UNSUPPORTED;
    }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(stem_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(stem_block) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    CUtensorMap m1, m2;
    int rc = out_map(a.y1, a.y1_ld, a, g.Wp, g.R, &m1);
    if (rc == SPK_OK) rc = out_map(a.y2, a.y2_ld, a, g.Wp, g.R, &m2);
    if (rc != SPK_OK) return rc;
    const long long items = (long long)a.B * g.n_bands;
    const long long grid = std::min<long long>(items, sm_count());
    const cudaError_t le = launch_pdl(stem_block_kernel, dim3((unsigned)grid), dim3(kThreads), (size_t)g.smem_bytes, s, a, g, items, m1, m2);
    if (le != cudaSuccess) {
        set_error("stem_block_kernel launch failed: %s", cudaGetErrorString(le));
        return SPK_ERR_CUDA;
    }
    return check_launch("stem_block_kernel");
}

}  // namespace spk

// debug aids (not part of the ABI)
extern "C" int spk_debug_stem_enable(int on) { return cudaMemcpyToSymbol(spk::g_stem_dbg, &on, sizeof(int)) == cudaSuccess ? 0 : -1; }
extern "C" int spk_debug_stem_timeline(long long *dst) {
    return cudaMemcpyFromSymbol(dst, spk::g_stem_ts, sizeof(long long) * 64 * 8) == cudaSuccess ? 0 : -1;
}
