// "Slab" implicit-GEMM for the small-channel 2-D convolutions of the CAM++ FCM head
// (BasicResBlock 3x3 convs, the 1x1 strided shortcuts and head.conv2:
// speakerlab/models/campplus/DTDNN.py:13-48, layers.py:218-253): Cin = Cout = 32, kernel 3x3
// pad 1 or 1x1 pad 0, stride (1|2, 1).  bf16 operands, fp32 accumulation in TMEM.
//
// Why not the generic gather kernel: with K = 9*32 every input pixel is needed by 9 taps, and a
// per-tap gather re-reads it 9 times through L1.  Here a CTA stages a band of input rows ONCE in
// shared memory as four 16-byte "channel planes" (plane j holds channels 8j..8j+7 of every
// pixel, pixels 16 B apart).  That is the no-swizzle K-major UMMA layout with SBO = 128 B
// (8 pixels) and LBO = plane stride, and - because consecutive pixels are a constant 16 B apart
// - the SAME staged data serves every tap: tap (kh,kw) is just the descriptor start address
// advanced by (kh*Wp + kw) pixels over the zero-padded, flattened band (Wp = W + KW - 1).
// For stride 2 in H, even and odd input rows go to separate sub-slabs so the shift stays uniform.
// Output pixels are produced for all Wp flat columns; the KW-1 wrap-around columns per row are
// computed and dropped (1.3 % at W=148).
//
// This is the cp.async generation of the kernel (warp-specialised, double buffered).  The default path for the
// CAM++ shapes is conv_slab3.cu (TMA-staged 64B-swizzled slabs, TMA store); this one remains for row pitches
// beyond one TMA box (W + 2 > 256 pixels: 3 s and 10 s segments, ERes2NetV2 layer 1) and as SPK_SLAB_V2=1 reference.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "ops.cuh"
#include "tc.cuh"

namespace spk {
namespace {

using namespace tc;        // mbarrier / TMEM / tcgen05 wrappers shared by all tensor-core kernels
using bf16 = __nv_bfloat16;
constexpr int kC = 32;                 // Cin = Cout


// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; SBO = next 8 rows, LBO = next
// 16-byte K chunk (cute::UMMA INTERLEAVE layout ((8,n),2):((1,SBO),LBO) in 16-byte units).
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
constexpr uint32_t kIdescN32 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);


struct SlabGeom {
    int R;          // output rows per band
    int Wp;         // staged row pitch in pixels: W + KW - 1 rounded up to a multiple of 8
    int n_bands;    // per segment
    int n_tiles;    // 128-pixel MMA tiles per band
    int rows_e, rows_o;     // input rows staged per sub-slab
    int px_e, px_o;         // pixels allocated per sub-slab (incl. slack)
    int taps;
    int smem_bytes;
    int tmem_cols;
    unsigned wp_magic;      // ceil(2^32 / Wp): exact pix / Wp for pix < 2^16
};

// S: stride in H (1|2); KS: kernel size (1|3)
// Warp-specialised and double buffered.  Two slab buffers and two
// TMEM accumulator sets let the three roles run one band apart:
//   warps 0-3   producers: 16-byte cp.async (LDGSTS, zero-fill for the conv padding) straight
//               into the channel planes - no register staging, so each thread has its ~33 copies
//               of a band in flight at once instead of paying the L2 latency per batch
//               (a TMA tiled load was tried first: with a 16-byte innermost box the TMA unit
//               processes one 16-byte line at a time and took ~10 us per band)
//   warp 4      MMA issuer (+ TMEM allocation)
//   warps 5-12  epilogue (two warps per TMEM lane quarter, alternating tiles)
constexpr int kProdThreads2 = 128;
constexpr int kEpiThreads2 = 256;
constexpr int kThreads2 = kProdThreads2 + 32 + kEpiThreads2;      // 416

// 16-byte async copy global -> shared; src_bytes = 0 writes zeros (conv padding)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int S, int KS>
__global__ void __launch_bounds__(kThreads2, 1)
conv_slab2_kernel(const ConvArgs a, const SlabGeom g, long long n_items) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int TAPS = KS * KS;
    constexpr int PAD = (KS - 1) / 2;
    const uint32_t s_w = smem_u32(smem);
    const uint32_t slab_bytes = 64u * (uint32_t)(g.px_e + g.px_o);
    const uint32_t s_slab0 = s_w + TAPS * 4 * 32 * 16;
    const uint32_t s_bar = s_slab0 + 2u * slab_bytes;
    // barriers: slab_full[2], slab_empty[2], acc_full[2], acc_empty[2]
    auto sfull = [&](int i) { return s_bar + 8u * i; };
    auto sempty = [&](int i) { return s_bar + 8u * (2 + i); };
    auto afull = [&](int i) { return s_bar + 8u * (4 + i); };
    auto aempty = [&](int i) { return s_bar + 8u * (6 + i); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + (s_bar - s_w) + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t plane_e = (uint32_t)g.px_e * 16u, plane_o = (uint32_t)g.px_o * 16u;
    const uint32_t acc_cols = (uint32_t)g.n_tiles * 32u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(sfull(i), kProdThreads2);
            mbar_init(sempty(i), 1);
            mbar_init(afull(i), 1);
            mbar_init(aempty(i), kEpiThreads2);
        }
        fence_barrier_init();
    }
    if (warp == 4) {
        __syncwarp();
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), g.tmem_cols);
    }
    {   // weights -> smem planes; slack pixels of both slabs -> 0 (TMA never writes them)
        const bf16 *w = static_cast<const bf16 *>(a.w);
        for (int idx = threadIdx.x; idx < kC * TAPS * 4; idx += kThreads2) {
            const int j = idx & 3, t = (idx >> 2) % TAPS, n = idx / (4 * TAPS);
            sts16(s_w + (uint32_t)(((t * 4 + j) * 32 + n) * 16), ldg16(w + ((long long)n * TAPS + t) * kC + j * 8));
        }
        const int slack_e = g.px_e - g.rows_e * g.Wp, slack_o = g.px_o - g.rows_o * g.Wp;
        for (int idx = threadIdx.x; idx < 2 * (slack_e + slack_o) * 4; idx += kThreads2) {
            const int j = idx & 3;
            int p = idx >> 2;
            const uint32_t sb = s_slab0 + (p >= slack_e + slack_o ? slab_bytes : 0u);
            if (p >= slack_e + slack_o) p -= slack_e + slack_o;
            const uint32_t dst = (p < slack_e) ? sb + j * plane_e + (uint32_t)(g.rows_e * g.Wp + p) * 16u
                                               : sb + 4u * plane_e + j * plane_o + (uint32_t)(g.rows_o * g.Wp + (p - slack_e)) * 16u;
            sts16(dst, make_uint4(0u, 0u, 0u, 0u));
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =========================== producers (cp.async) ===========================
        const bf16 *x = static_cast<const bf16 *>(a.x);
        uint32_t it = 0;
        const int rows_total = g.rows_e + g.rows_o;
        const int pieces = rows_total * g.Wp * 4;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b = (int)(item / g.n_bands);
            const int band = (int)(item - (long long)b * g.n_bands);
            const int hi_base = band * g.R * S - PAD;
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(sempty(buf), ph ^ 1u);
            const uint32_t se = s_slab0 + buf * slab_bytes, so = se + 4u * plane_e;
            for (int idx = threadIdx.x; idx < pieces; idx += kProdThreads2) {
                const int j = idx & 3;
                const int pix = idx >> 2;
                const int row = (int)__umulhi((unsigned)pix, g.wp_magic);
                const int col = pix - row * g.Wp;
                int sub = 0, srow = row, hi;
                if (S == 2 && KS == 3) {
                    if (row < g.rows_e) { sub = 0; srow = row; hi = hi_base + 2 * row; }
                    else { sub = 1; srow = row - g.rows_e; hi = hi_base + 2 * srow + 1; }
                } else if (S == 2) {
                    hi = hi_base + 2 * row;
                } else {
                    hi = hi_base + row;
                }
                const int wi = col - PAD;
                const bool ok = hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
                const bf16 *src = ok ? x + (((long long)b * a.H + hi) * a.W + wi) * a.in_ld + a.in_choff + j * 8 : x;
                const uint32_t dst = (sub == 0 ? se + j * plane_e : so + j * plane_o) + (uint32_t)(srow * g.Wp + col) * 16u;
                cp_async16(dst, src, ok ? 16u : 0u);
            }
            cp_async_wait_all();
            fence_proxy_async();
            mbar_arrive(sfull(buf));
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        uint32_t it = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(aempty(buf), ph ^ 1u);
            mbar_wait(sfull(buf), ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t se = s_slab0 + buf * slab_bytes, so = se + 4u * plane_e;
                for (int t = 0; t < g.n_tiles; ++t) {
                    const uint32_t d = tmem_base + buf * acc_cols + (uint32_t)t * 32u;
#pragma unroll
                    for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                        for (int kw = 0; kw < KS; ++kw) {
                            uint32_t sbase, plane;
                            int off;
                            if (S == 2 && KS == 3) {
                                if (kh & 1) { sbase = so; plane = plane_o; off = kw; }
                                else { sbase = se; plane = plane_e; off = (kh >> 1) * g.Wp + kw; }
                            } else {
                                sbase = se; plane = plane_e; off = kh * g.Wp + kw;
                            }
                            const uint32_t a_addr = sbase + (uint32_t)(t * 128 + off) * 16u;
                            const uint32_t b_addr = s_w + (uint32_t)((kh * KS + kw) * 4 * 32 * 16);
#pragma unroll
                            for (int half = 0; half < 2; ++half)
                                umma_bf16(d, make_desc_nosw(a_addr + 2u * half * plane, plane, 128u),
                                          make_desc_nosw(b_addr + 2u * half * 512u, 512u, 128u), kIdescN32,
                                          (kh | kw | half) ? 1u : 0u);
                        }
                }
                umma_commit(sempty(buf));      // slab reusable once these MMAs retire
                umma_commit(afull(buf));
            }
            __syncwarp();
        }
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;
        const int tsel = (warp - 5) >> 2;           // 0 or 1: even / odd tiles
        bf16 *y = static_cast<bf16 *>(a.y);
        const bf16 *res = static_cast<const bf16 *>(a.res);
        uint32_t it = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b = (int)(item / g.n_bands);
            const int band = (int)(item - (long long)b * g.n_bands);
            const int ho0 = band * g.R;
            const int r_valid = min(g.R, a.Ho - ho0);
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            bool waited = false;
            for (int t = tsel; t < g.n_tiles; t += 2) {
                const int p = t * 128 + q * 32 + lane;
                const int i = (int)__umulhi((unsigned)p, g.wp_magic), col = p - i * g.Wp;
                const bool ok = (i < r_valid) && (col < a.W);
                const long long opix = ((long long)b * a.Ho + ho0 + i) * a.Wo + col;
                uint4 rr4[4];
                if (res != nullptr && ok) {      // independent of the MMAs: issue before waiting
#pragma unroll
                    for (int e = 0; e < 4; ++e) rr4[e] = ldg16(res + opix * a.res_ld + a.res_choff + e * 8);
                }
                if (!waited) {
                    mbar_wait(afull(buf), ph);
                    tc_fence_after();
                    waited = true;
                }
                const uint32_t taddr = tmem_base + buf * acc_cols + (uint32_t)t * 32u + ((uint32_t)(q * 32) << 16);
                uint32_t r[32];
                {
                    uint32_t (&r0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[0]);
                    uint32_t (&r1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[16]);
                    tmem_ld16(taddr, r0);
                    tmem_ld16(taddr + 16, r1);
                    tmem_ld_wait();
                }
                if (ok) {
                    float v[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
                    if (a.epi_scale != nullptr) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            const float4 s4 = __ldg(reinterpret_cast<const float4 *>(a.epi_scale + e));
                            const float4 h4 = __ldg(reinterpret_cast<const float4 *>(a.epi_shift + e));
                            v[e] = fmaf(v[e], s4.x, h4.x); v[e + 1] = fmaf(v[e + 1], s4.y, h4.y);
                            v[e + 2] = fmaf(v[e + 2], s4.z, h4.z); v[e + 3] = fmaf(v[e + 3], s4.w, h4.w);
                        }
                    }
                    if (res != nullptr) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t w4[4] = {rr4[e].x, rr4[e].y, rr4[e].z, rr4[e].w};
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const float2 f = unpack2(w4[h]);
                                v[e * 8 + 2 * h] += f.x;
                                v[e * 8 + 2 * h + 1] += f.y;
                            }
                        }
                    }
                    apply_act_vec(v, a.act);
                    bf16 *yp = y + opix * a.out_ld + a.out_choff;
#pragma unroll
                    for (int e = 0; e < 32; e += 8)
                        *reinterpret_cast<uint4 *>(yp + e) =
                            make_uint4(pack2(v[e], v[e + 1]), pack2(v[e + 2], v[e + 3]), pack2(v[e + 4], v[e + 5]), pack2(v[e + 6], v[e + 7]));
                }
            }
            if (!waited) {                      // this warp had no tile in the band: still keep the phase in step
                mbar_wait(afull(buf), ph);
                tc_fence_after();
            }
            tc_fence_before();
            mbar_arrive(aempty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, g.tmem_cols);
}

bool geometry(const ConvArgs &a, SlabGeom &g, bool v2 = false) {
    const int KS = a.KH;
    g.taps = KS * KS;
    // row pitch of the staged band in pixels: W + KS - 1 rounded up to 8 pixels so that every slab
    // row starts 128-byte aligned (TMA destination alignment); the extra columns are zero (OOB)
    g.Wp = (a.W + KS - 1 + 7) & ~7;
    int R = 768 / g.Wp;
    if (R < 1) R = 1;
    if (R > a.Ho) R = a.Ho;
    const int smem_cap = v2 ? 210 * 1024 : 110 * 1024;
    for (;; --R) {
        g.R = R;
        g.n_tiles = (R * g.Wp + 127) / 128;
        if (a.sh == 1) { g.rows_e = R + KS - 1; g.rows_o = 0; }
        else if (KS == 3) { g.rows_e = R + 1; g.rows_o = R; }
        else { g.rows_e = R; g.rows_o = 0; }
        const int max_off_e = (a.sh == 1 ? (KS - 1) * g.Wp : (KS == 3 ? g.Wp : 0)) + (KS - 1);
        g.px_e = (std::max(g.rows_e * g.Wp, g.n_tiles * 128 + max_off_e) + 8 + 7) & ~7;
        g.px_o = g.rows_o ? (std::max(g.rows_o * g.Wp, g.n_tiles * 128 + (KS - 1)) + 8 + 7) & ~7 : 0;
        const int slab = 64 * (g.px_e + g.px_o);
        g.smem_bytes = g.taps * 4 * 32 * 16 + (v2 ? 2 : 1) * slab + 128;
        const int cols = g.n_tiles * 32 * (v2 ? 2 : 1);
        g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
        if ((cols <= (v2 ? 512 : 256) && g.smem_bytes <= smem_cap) || R == 1) break;
    }
    g.n_bands = (a.Ho + g.R - 1) / g.R;
    g.wp_magic = (unsigned)(((1ull << 32) + g.Wp - 1) / g.Wp);
    if ((long long)(g.rows_e + g.rows_o) * g.Wp >= 65536 || g.n_tiles * 128 >= 65536) return false;
    return g.n_tiles * 32 * (v2 ? 2 : 1) <= 512 && g.smem_bytes <= (v2 ? 220 : 200) * 1024;
}

template <int S, int KS>
int launch2(const ConvArgs &a, const SlabGeom &g, cudaStream_t s) {
    auto kern = conv_slab2_kernel<S, KS>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_slab2) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    const long long items = (long long)a.B * g.n_bands;
    const long long grid = std::min<long long>(items, sm_count());
    kern<<<(unsigned)grid, kThreads2, g.smem_bytes, s>>>(a, g, items);
    return check_launch("conv_slab2_kernel");
}

}  // namespace

bool conv_slab_supported(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype) {
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
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0) return false;
    SlabGeom g;
    return geometry(a, g, true);
}

int launch_conv_slab(const ConvArgs &a, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    SlabGeom g;
    if (!geometry(a, g, true)) {
        set_error("conv_slab: geometry does not fit");
        return SPK_ERR_UNSUPPORTED;
    }
    if (a.KH == 3) return a.sh == 1 ? launch2<1, 3>(a, g, s) : launch2<2, 3>(a, g, s);
    return launch2<2, 1>(a, g, s);
}

}  // namespace spk
