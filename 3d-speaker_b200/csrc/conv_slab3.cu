// Third-generation slab kernel for the Cin = Cout = 32 2-D convs of the CAM++ head (FCM): 3x3 stride (1|2, 1)
// and the 1x1 stride (2, 1) shortcut (speakerlab/models/campplus/DTDNN.py:19-63).
//
// An item is a band of R output rows of one segment.  Its input rows are staged ONCE in shared memory by a single
// TMA box {32 channels, Wp pixels, rows, 1 segment} that starts at (w = -1, h = h0 - 1): the conv zero padding - left,
// right, top, bottom - is the TMA unit's out-of-bounds fill, and a stride-2 conv reads every other row through the
// tensor map's traversal stride (even and odd input rows land in two sub-slabs).  The slab is K-major with 64-byte
// rows (one pixel = 32 bf16 channels) and the 64-byte swizzle TMA writes natively.  A tap (kh, kw) of the conv is
// then the same slab viewed through a UMMA descriptor whose start address is shifted by kh*Wp + kw pixels - the
// swizzle is a function of the absolute shared-memory address, so any pixel shift keeps the pattern - and an
// output tile of 128 consecutive slab pixels is 2*KH*KW accumulating tcgen05.mma (M=128, N=32, K=16).
//
// Small-N MMAs are bound by the shared-memory read of A (4 KB each, ~40 cycles) and by the issue rate of the one
// issuing thread; the issue loop is fully unrolled with every operand one add away (measured: tools/mma_*_bench.cu).
//
// The output leaves the same way it came in.  Per-lane 16-byte global stores of a 64-byte pixel (and the matching
// residual loads) touch sixteen 128-byte lines per warp instruction and made the epilogue the bottleneck (measured:
// 214 -> 80 us per launch with the stores removed).  Instead the epilogue writes the band into a 64B-swizzled staging
// buffer whose pixel order IS the slab pixel order, and one TMA box store {32 ch, Wp px, R rows} writes it out; pad
// columns and rows past the image are clipped by the tensor map.  The residual band is TMA-loaded into the same
// staging buffer beforehand and updated in place.
//
//   warp 0      TMA producer (one elected thread: expect_tx + 1-2 slab boxes + the residual box per item)
//   warp 1      MMA issuer (+ TMEM allocation)
//   warps 2-9   epilogue, two warps per TMEM lane quarter on alternating tiles: BN scale/shift, residual, activation,
//               bf16 into the staging buffer
//   warp 10     TMA store of the finished band
// Two TMEM accumulator sets keep MMA and epilogue one item apart; the slab ring is as deep as shared memory allows
// (2-4 buffers): with ~70 KB items the bytes in flight per SM, not the copy engine, set the HBM rate.
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
constexpr int kEpi = 256, kThreads = 64 + kEpi + 32;  // 352
constexpr uint32_t kIdesc = idesc_bf16(32);

struct Slab3Geom {
    int Wp, R, n_tiles, n_bands;
    int rows_e, rows_o;       // slab rows of the (even) and odd sub-slab
    int px_e, px_o;           // pixels per sub-slab including the slack the shifted views may touch
    int smem_bytes, tmem_cols;
    int nbuf;                 // slab buffers (2-4): bytes in flight per SM = (nbuf - 1) slabs, which is what sets the HBM rate
    unsigned wp_magic;        // ceil(2^32 / Wp)
    uint32_t off_w, off_slab, slab_bytes, off_stg, stg_bytes, off_bar;
};

// debug aid: per-item role timestamps of CTA 0 (SPK_SLAB_DBG=1), read back by spk_debug_slab_timeline
__device__ long long g_slab_ts[64 * 8];
#define SLAB_TS(idx, slot) do { if (dbg && blockIdx.x == 0 && (idx) < 64 && (threadIdx.x & 31) == 0) g_slab_ts[(idx) * 8 + (slot)] = clock64(); } while (0)

template <int S, int KS>
__global__ void __launch_bounds__(kThreads, 1)
conv_slab3_kernel(const ConvArgs a, const Slab3Geom g, long long n_items, const __grid_constant__ CUtensorMap xmap_e,
                  const __grid_constant__ CUtensorMap xmap_o, const __grid_constant__ CUtensorMap ymap,
                  const __grid_constant__ CUtensorMap rmap, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int TAPS = KS * KS;
    constexpr int PAD = (KS - 1) / 2;
    const uint32_t s0 = smem_u32(smem);
    const uint32_t s_w = s0 + g.off_w, s_slab0 = s0 + g.off_slab, s_bar = s0 + g.off_bar;
    auto sfull = [&](int i) { return s_bar + 8u * i; };
    auto sempty = [&](int i) { return s_bar + 8u * (4 + i); };
    auto afull = [&](int i) { return s_bar + 8u * (8 + i); };
    auto aempty = [&](int i) { return s_bar + 8u * (10 + i); };
    auto rfull = [&](int i) { return s_bar + 8u * (12 + i); };      // residual band landed in the staging buffer
    auto gfull = [&](int i) { return s_bar + 8u * (14 + i); };      // staging buffer holds the finished band
    auto gfree = [&](int i) { return s_bar + 8u * (16 + i); };      // the TMA store has read the staging buffer
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + g.off_bar + 160);
    const uint32_t s_stg = s0 + g.off_stg;
    const bool has_res = a.res != nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t acc_cols = (uint32_t)g.n_tiles * 32u;
    const uint32_t sub_o = (uint32_t)g.px_e * 64u;          // byte offset of the odd sub-slab inside a slab buffer

    pdl_trigger();
    if (threadIdx.x == 0) {
        tmap_prefetch(&xmap_e);
        if (S == 2 && KS == 3) tmap_prefetch(&xmap_o);
        for (int i = 0; i < 4; ++i) {
            mbar_init(sfull(i), 1);
            mbar_init(sempty(i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(afull(i), 1);
            mbar_init(aempty(i), kEpi);
            mbar_init(rfull(i), 1);
            mbar_init(gfull(i), kEpi);
            mbar_init(gfree(i), 1);
        }
        tmap_prefetch(&ymap);
        if (has_res) tmap_prefetch(&rmap);
        fence_barrier_init();
    }
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), g.tmem_cols);
    }
    {   // weights -> smem: per tap a K-major 64B-swizzled [32 cout][32 cin] block (16-byte chunk c of row n at
        // c ^ ((n >> 1) & 3)); pixels behind the last staged row of each sub-slab -> 0 (TMA never writes them)
        const bf16 *w = static_cast<const bf16 *>(a.w);      // [32][TAPS][32]
        for (int idx = threadIdx.x; idx < kC * TAPS * 4; idx += kThreads) {
            const int c = idx & 3, t = (idx >> 2) % TAPS, n = idx / (4 * TAPS);
            sts16(s_w + (uint32_t)(t * 2048 + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)), ldg16(w + ((long long)n * TAPS + t) * kC + c * 8));
        }
        const int slack_e = g.px_e - g.rows_e * g.Wp, slack_o = g.px_o - g.rows_o * g.Wp;
        for (int idx = threadIdx.x; idx < g.nbuf * (slack_e + slack_o) * 4; idx += kThreads) {
            const int c = idx & 3;
            int p = idx >> 2;
            const int bi = p / (slack_e + slack_o);
            const uint32_t sb = s_slab0 + (uint32_t)bi * g.slab_bytes;
            p -= bi * (slack_e + slack_o);
            const uint32_t dst = p < slack_e ? sb + (uint32_t)(g.rows_e * g.Wp + p) * 64u
                                             : sb + sub_o + (uint32_t)(g.rows_o * g.Wp + (p - slack_e)) * 64u;
            sts16(dst + (uint32_t)c * 16u, make_uint4(0u, 0u, 0u, 0u));
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
            const uint32_t bytes = 64u * (uint32_t)((g.rows_e + g.rows_o) * g.Wp);
            uint32_t buf = 0, ph = 0, it = 0;
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b = (int)(item / g.n_bands);
                const int band = (int)(item - (long long)b * g.n_bands);
                const int hi0 = band * g.R * S - PAD;           // first input row of the band
                mbar_wait(sempty(buf), ph ^ 1u);
                SLAB_TS(it, 0);
                mbar_arrive_expect_tx(sfull(buf), bytes);
                const uint32_t sb = s_slab0 + buf * g.slab_bytes;
                tmap_load_4d(sb, &xmap_e, a.in_choff, -PAD, hi0, b, sfull(buf));
                if (S == 2 && KS == 3) tmap_load_4d(sb + sub_o, &xmap_o, a.in_choff, -PAD, hi0 + 1, b, sfull(buf));
                if (has_res) {      // residual band -> staging buffer, once the store that last used it has read it
                    const uint32_t gb = it & 1u, gph = (it >> 1) & 1u;
                    mbar_wait(gfree(gb), gph ^ 1u);
                    mbar_arrive_expect_tx(rfull(gb), 64u * (uint32_t)(g.R * g.Wp));
                    tmap_load_4d(s_stg + gb * g.stg_bytes, &rmap, a.res_choff, 0, band * g.R, b, rfull(gb));
                }
                if (++buf == (uint32_t)g.nbuf) { buf = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        const uint32_t hi = desc_hi(512u, kLayoutSw64);         // 8-pixel groups are 512 B apart
        const uint32_t wp4 = (uint32_t)g.Wp * 4u;               // one slab row, in 16-byte descriptor units
        uint32_t it = 0, sbuf = 0, sph = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(aempty(buf), ph ^ 1u);
            SLAB_TS(it, 1);
            mbar_wait(sfull(sbuf), sph);
            tc_fence_after();
            SLAB_TS(it, 2);
            if (elect_one()) {
                const uint32_t lo_e = desc_lo(s_slab0 + sbuf * g.slab_bytes, 16u), lo_o = lo_e + (sub_o >> 4);
                const uint32_t lo_w = desc_lo(s_w, 16u);
                uint32_t d = tmem_base + buf * acc_cols;
                uint32_t tile = 0;                                // 128 pixels = 8192 B = 512 units
                for (int t = 0; t < g.n_tiles; ++t, d += 32u, tile += 512u) {
#pragma unroll
                    for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                        for (int kw = 0; kw < KS; ++kw) {
                            // slab pixel shift of this tap, in 16-byte units (4 per pixel)
                            uint32_t lo_a;
                            if (S == 2 && KS == 3) lo_a = (kh & 1) ? lo_o + tile + 4u * kw : lo_e + tile + (kh >> 1) * wp4 + 4u * kw;
                            else lo_a = lo_e + tile + kh * wp4 + 4u * kw;
                            const uint32_t lo_b = lo_w + (uint32_t)((kh * KS + kw) * (2048 >> 4));
                            if (kh == 0 && kw == 0) umma_bf16(d, desc64(lo_a, hi), desc64(lo_b, hi), kIdesc, 0u);
                            else umma_bf16_acc(d, desc64(lo_a, hi), desc64(lo_b, hi), kIdesc);
                            umma_bf16_acc(d, desc64(lo_a + 2u, hi), desc64(lo_b + 2u, hi), kIdesc);       // channels 16-31
                        }
                }
                umma_commit(sempty(sbuf));     // slab reusable once these MMAs retire
                umma_commit(afull(buf));
            }
            __syncwarp();
            SLAB_TS(it, 3);
            if (++sbuf == (uint32_t)g.nbuf) { sbuf = 0; sph ^= 1u; }
        }
    } else if (warp < 10) {
        // =========================== epilogue ===========================
        const int q = warp & 3;
        const int tsel = (warp - 2) >> 2;           // 0 or 1: even / odd tiles
        float sc[32], sh[32];
        if (a.epi_scale != nullptr) {
#pragma unroll
            for (int e = 0; e < 32; ++e) { sc[e] = __ldg(a.epi_scale + e); sh[e] = __ldg(a.epi_shift + e); }
        }
        uint32_t it = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            const uint32_t stg = s_stg + buf * g.stg_bytes;
            // the staging buffer is ours once the residual has landed in it (which implies the previous store has read
            // it) or, without a residual, once that store has read it
            if (has_res) mbar_wait(rfull(buf), ph);
            else mbar_wait(gfree(buf), ph ^ 1u);
            mbar_wait(afull(buf), ph);
            tc_fence_after();
            if (warp == 2) SLAB_TS(it, 4);
            for (int t = tsel; t < g.n_tiles; t += 2) {
                const int p = t * 128 + q * 32 + lane;               // slab pixel == staging pixel
                const uint32_t taddr = tmem_base + buf * acc_cols + (uint32_t)t * 32u + ((uint32_t)(q * 32) << 16);
                uint32_t r[32];
                {
                    uint32_t (&r0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[0]);
                    uint32_t (&r1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[16]);
                    tmem_ld16(taddr, r0);
                    tmem_ld16(taddr + 16, r1);
                    tmem_ld_wait();
                }
                if (warp == 2 && t == tsel) SLAB_TS(it, 6);
                float v[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
                if (a.epi_scale != nullptr) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
                }
                const uint32_t prow = stg + (uint32_t)p * 64u;
                const uint32_t x = (uint32_t)(p >> 1) & 3u;         // 64B swizzle: 16-byte chunk c lives at c ^ x
                if (has_res) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint4 w = lds16(prow + (((uint32_t)e ^ x) << 4));
                        const uint32_t w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 f = unpack2(w4[h]);
                            v[e * 8 + 2 * h] += f.x;
                            v[e * 8 + 2 * h + 1] += f.y;
                        }
                    }
                }
                apply_act_vec(v, a.act);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    sts16(prow + (((uint32_t)e ^ x) << 4),
                          make_uint4(pack2(v[8 * e], v[8 * e + 1]), pack2(v[8 * e + 2], v[8 * e + 3]), pack2(v[8 * e + 4], v[8 * e + 5]),
                                     pack2(v[8 * e + 6], v[8 * e + 7])));
                if (warp == 2 && t == tsel) SLAB_TS(it, 7);
            }
            tc_fence_before();
            mbar_arrive(aempty(buf));
            fence_proxy_async();            // staging writes (generic proxy) -> TMA store (async proxy)
            mbar_arrive(gfull(buf));
            if (warp == 2) SLAB_TS(it, 5);
        }
    } else {
        // =========================== TMA store ===========================
        if (elect_one()) {
            uint32_t it = 0;
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b = (int)(item / g.n_bands);
                const int band = (int)(item - (long long)b * g.n_bands);
                const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
                mbar_wait(gfull(buf), ph);
                tmap_store_4d(&ymap, a.out_choff, 0, band * g.R, b, s_stg + buf * g.stg_bytes);
                bulk_commit();
                bulk_wait_read0();          // the box has been read out of shared memory
                mbar_arrive(gfree(buf));
            }
            bulk_wait_all();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, g.tmem_cols);
}

bool geometry(const ConvArgs &a, Slab3Geom &g) {
    const int KS = a.KH;
    // row pitch of the staged band in pixels: W + KS - 1 rounded up to the 8-pixel swizzle atom; the extra
    // columns are out-of-bounds zeros
    g.Wp = (a.W + KS - 1 + 7) & ~7;
    if (g.Wp > 256) return false;               // one TMA box per sub-slab
    int R = 768 / g.Wp;
    if (R < 1) R = 1;
    if (R > a.Ho) R = a.Ho;
    for (;; --R) {
        g.R = R;
        g.n_tiles = (R * g.Wp + 127) / 128;
        if (a.sh == 1) { g.rows_e = R + KS - 1; g.rows_o = 0; }
        else if (KS == 3) { g.rows_e = R + 1; g.rows_o = R; }
        else { g.rows_e = R; g.rows_o = 0; }
        const int max_off_e = (a.sh == 1 ? (KS - 1) * g.Wp : (KS == 3 ? g.Wp : 0)) + (KS - 1);
        g.px_e = (std::max(g.rows_e * g.Wp, g.n_tiles * 128 + max_off_e) + 7) & ~7;
        g.px_o = g.rows_o ? (std::max(g.rows_o * g.Wp, g.n_tiles * 128 + (KS - 1)) + 7) & ~7 : 0;
        g.slab_bytes = (64u * (uint32_t)(g.px_e + g.px_o) + 1023u) & ~1023u;
        g.stg_bytes = (uint32_t)g.n_tiles * 128u * 64u;          // every tile row has a slot; the store box reads R*Wp of them
        g.off_w = 0;
        g.off_stg = (uint32_t)((KS * KS * 2048 + 1023) & ~1023);
        g.off_slab = g.off_stg + 2u * g.stg_bytes;
        g.nbuf = 2;
        g.off_bar = g.off_slab + 2u * g.slab_bytes;
        g.smem_bytes = (int)g.off_bar + 256;
        const int cols = g.n_tiles * 32 * 2;
        g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
        if ((cols <= 512 && g.smem_bytes <= 226 * 1024) || R == 1) break;
    }
    // spend what is left of shared memory on more slabs in flight
    while (g.nbuf < 4 && g.off_slab + (uint32_t)(g.nbuf + 1) * g.slab_bytes + 256u <= 226u * 1024u) ++g.nbuf;
    g.off_bar = g.off_slab + (uint32_t)g.nbuf * g.slab_bytes;
    g.smem_bytes = (int)g.off_bar + 256;
    g.n_bands = (a.Ho + g.R - 1) / g.R;
    g.wp_magic = (unsigned)(((1ull << 32) + g.Wp - 1) / g.Wp);
    // TMA box limits: rows (times the traversal stride) at most 256
    if (g.rows_e * a.sh > 256 || g.rows_o * a.sh > 256 || g.R > 256) return false;
    return g.n_tiles * 64 <= 512 && g.smem_bytes <= 227 * 1024;
}

// {32 channels, Wp pixels, rows (every sh-th), 1 segment} boxes over the [B][H][W][ld] activation buffer, 64B swizzle
int input_map(const ConvArgs &a, int wp, int rows, CUtensorMap *out) {
    typedef std::tuple<const void *, int, int, int, int, int, int, int> Key;
    static std::mutex mu;
    static std::map<Key, CUtensorMap> cache;
    const Key key(a.x, a.in_ld, a.W, a.H, a.B, wp, rows, a.sh);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return SPK_OK;
    }
    TmapEncodeTiledFn fn = tmap_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SPK_ERR_CUDA;
    }
    const cuuint64_t dims[4] = {(cuuint64_t)a.in_ld, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    const cuuint64_t strides[3] = {(cuuint64_t)a.in_ld * 2, (cuuint64_t)a.W * a.in_ld * 2, (cuuint64_t)a.H * a.W * a.in_ld * 2};
    const cuuint32_t box[4] = {32u, (cuuint32_t)wp, (cuuint32_t)(rows * a.sh), 1u};
    const cuuint32_t estr[4] = {1u, 1u, (cuuint32_t)a.sh, 1u};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(a.x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): slab box {32,%d,%d} stride %d over [%d][%d][%d][%d]", (int)r, wp, rows, a.sh,
                  a.B, a.H, a.W, a.in_ld);
        return SPK_ERR_CUDA;
    }
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return SPK_OK;
}

// {32 channels, Wp pixels, R rows, 1 segment} boxes over a [B][Ho][Wo][ld] buffer (the output, or the residual), 64B swizzle
int band_map(const void *ptr, int ld, const ConvArgs &a, int wp, int rows, CUtensorMap *out) {
    typedef std::tuple<const void *, int, int, int, int, int, int> Key;
    static std::mutex mu;
    static std::map<Key, CUtensorMap> cache;
    const Key key(ptr, ld, a.Wo, a.Ho, a.B, wp, rows);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return SPK_OK;
    }
    const uint64_t dims[4] = {(uint64_t)ld, (uint64_t)a.Wo, (uint64_t)a.Ho, (uint64_t)a.B};
    const uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)a.Wo * ld * 2, (uint64_t)a.Ho * a.Wo * ld * 2};
    const uint32_t box[4] = {32u, (uint32_t)wp, (uint32_t)rows, 1u};
    const int rc = tmap_encode_bf16(ptr, 4, dims, strides, box, 64, out);
    if (rc != SPK_OK) return rc;
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return SPK_OK;
}

template <int S, int KS>
int launch(const ConvArgs &a, const Slab3Geom &g, cudaStream_t s) {
    auto kern = conv_slab3_kernel<S, KS>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_slab3) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    CUtensorMap me, mo;
    int rc = input_map(a, g.Wp, g.rows_e, &me);
    if (rc != SPK_OK) return rc;
    mo = me;
    if (g.rows_o) {
        rc = input_map(a, g.Wp, g.rows_o, &mo);
        if (rc != SPK_OK) return rc;
    }
    CUtensorMap my, mr;
    rc = band_map(a.y, a.out_ld, a, g.Wp, g.R, &my);
    if (rc != SPK_OK) return rc;
    mr = my;
    if (a.res != nullptr) {
        rc = band_map(a.res, a.res_ld, a, g.Wp, g.R, &mr);
        if (rc != SPK_OK) return rc;
    }
    const long long items = (long long)a.B * g.n_bands;
    const long long grid = std::min<long long>(items, sm_count());
    static const int dbg = getenv("SPK_SLAB_DBG") ? atoi(getenv("SPK_SLAB_DBG")) : 0;
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), (size_t)g.smem_bytes, s, a, g, items, me, mo, my, mr, dbg);
    if (le != cudaSuccess) {
        set_error("conv_slab3_kernel launch failed: %s", cudaGetErrorString(le));
        return SPK_ERR_CUDA;
    }
    return check_launch("conv_slab3_kernel");
}

}  // namespace

bool conv_slab3_supported(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype) {
    if (in_dtype != SPK_DT_BF16 || out_dtype != SPK_DT_BF16) return false;
    if (a.res != nullptr && res_dtype != SPK_DT_BF16) return false;
    if (a.Cin != kC || a.Cout != kC) return false;
    if (!((a.KH == 3 && a.KW == 3 && a.ph == 1 && a.pw == 1) || (a.KH == 1 && a.KW == 1 && a.ph == 0 && a.pw == 0))) return false;
    if (a.sw != 1 || (a.sh != 1 && a.sh != 2) || a.dh != 1 || a.dw != 1) return false;
    if (a.pro_scale != nullptr || a.gate != nullptr) return false;
    if (a.post_scale != nullptr || a.pad_reflect) return false;
    if (a.in_ld % 8 || a.in_choff % 8 || a.out_ld % 8 || a.out_choff % 8) return false;
    if (a.res != nullptr && (a.res_ld % 8 || a.res_choff % 8)) return false;
    if (a.Wo != a.W || a.Ho != (a.H + 2 * a.ph - a.KH) / a.sh + 1) return false;
    if (a.KH == 1 && a.sh == 1) return false;      // plain 1x1: the generic GEMM path is already ideal
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.y) & 15) != 0) return false;
    if (a.res != nullptr && (reinterpret_cast<uintptr_t>(a.res) & 15) != 0) return false;
    Slab3Geom g;
    return geometry(a, g);
}

int launch_conv_slab3(const ConvArgs &a, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    Slab3Geom g;
    if (!geometry(a, g)) {
        set_error("conv_slab3: geometry does not fit");
        return SPK_ERR_UNSUPPORTED;
    }
    if (a.KH == 3) return a.sh == 1 ? launch<1, 3>(a, g, s) : launch<2, 3>(a, g, s);
    return launch<2, 1>(a, g, s);
}

}  // namespace spk

// debug aid (not part of the ABI)
extern "C" int spk_debug_slab_timeline(long long *dst) {
    return cudaMemcpyFromSymbol(dst, spk::g_slab_ts, sizeof(long long) * 64 * 8) == cudaSuccess ? 0 : -1;
}
