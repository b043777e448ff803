// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate).
//
//   D[m, n] = sum_k A[m, k] * W[n, k]        m = output pixel (b, ho, wo), n = output channel,
//                                            k = (kh, kw, cin) of the channels-last input
//
// One kernel serves every conv of the embedding networks (1x1 with BN-ReLU prologue on a
// growing concat buffer, dilated k=3 with a CAM gate epilogue, 3x3 / strided 2-D convs with
// residual epilogue, the KH=10 x KW=5 "tdnn" conv): the A operand is GATHERED by the producer
// warps (implicit im2col: zero padding, stride, dilation, optional per-channel scale/shift/ReLU
// prologue applied in registers) straight into the 128-byte-swizzled K-major shared-memory
// layout the tensor core reads, so no im2col buffer ever exists in HBM.
//
// Roles (288 threads, persistent over output tiles, static round-robin):
//   warps 0-3  producers: global -> registers (-> prologue) -> swizzled smem stage, then
//              fence.proxy.async + mbarrier arrive                        [full barrier]
//   warp  4    TMEM allocator + MMA issuer: one lane issues tcgen05.mma (M=128, N=BLOCK_N,
//              K=16) x4 per 64-wide K chunk; tcgen05.commit releases the smem stage
//              [empty barrier] and, after the last chunk, publishes the accumulator
//   warps 5-8  epilogue: tcgen05.ld the fp32 accumulator (one TMEM lane = one output pixel per
//              thread), per-channel affine (folded BN), residual add, activation, CAM gate,
//              convert and store channels-last.
// The accumulator is double buffered in TMEM (2 x BLOCK_N columns), so the epilogue of tile i
// overlaps the gather + MMA of tile i+1.
#include <cstdlib>
#include <mutex>

#include "ops.cuh"

namespace spk {
namespace {

using bf16 = __nv_bfloat16;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kProducerThreads = 128;
constexpr int kEpilogueThreads = 128;
constexpr int kThreads = kProducerThreads + 32 + kEpilogueThreads;
constexpr uint32_t kSpinLimit = 1u << 26;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must crash the context, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 B (64 bf16), 8-row groups of 1024 B.
// (cute::UMMA::SmemDescriptor: start>>4 | LBO[16,30) | SBO[32,46) | version=1 @46 | layout @61)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                      // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;            // SBO: next 8-row group
    d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ uint4 ldg16(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ float2 unpack2(uint32_t v) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162 *>(&v);
    return make_float2(__low2float(t), __high2float(t));
}

template <int BLOCK_N> struct Cfg {
    static constexpr int kStages = (BLOCK_N >= 128) ? 3 : 4;   // <=128: two CTAs per SM fit
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : 2 * BLOCK_N;     // 64..512, power of two
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BLOCK_N, typename TOut, typename TRes>
__global__ void __launch_bounds__(kThreads, (BLOCK_N >= 256) ? 1 : 2)
conv_tc_kernel(const ConvArgs a, int n_tiles_n, long long n_tiles) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *bar_ptr = smem_raw + (base - raw) + C::kStages * C::kStageBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bar_ptr);
    // [0,S) full  [S,2S) empty  [2S,2S+2) accum_full  [2S+2,2S+4) accum_empty ; then tmem ptr
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
    auto accf_bar = [&](int b) { return bar0 + 8u * (2 * C::kStages + b); };
    auto acce_bar = [&](int b) { return bar0 + 8u * (2 * C::kStages + 2 + b); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bars + 2 * C::kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(full_bar(s), kProducerThreads);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accf_bar(b), 1);
            mbar_init(acce_bar(b), kEpilogueThreads);
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nk = (a.K + BLOCK_K - 1) / BLOCK_K;
    const int HoWo = a.Ho * a.Wo;

    if (warp < 4) {
        // =========================== producers ===========================
        const int t = threadIdx.x;                 // 0..127
        const int j = t & 7;                       // 16-byte piece inside the 128-byte row
        const int r0 = t >> 3;                     // first row; rows r0 + 16*i
        const uint32_t sw_off = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
        const bf16 *x = static_cast<const bf16 *>(a.x);
        const bf16 *w = static_cast<const bf16 *>(a.w);
        int stage = 0;
        uint32_t phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long mt = tile / n_tiles_n;
            const int nt = (int)(tile - mt * n_tiles_n);
            const long long m0 = mt * BLOCK_M;
            const int n0 = nt * BLOCK_N;
            // decode this thread's 8 rows
            int pix0[8], hi0[8], wi0[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long m = m0 + r0 + 16 * i;
                if (m < a.M) {
                    const int b = (int)(m / HoWo);
                    const int r = (int)(m - (long long)b * HoWo);
                    const int ho = r / a.Wo, wo = r - ho * a.Wo;
                    pix0[i] = b * a.H * a.W;
                    hi0[i] = ho * a.sh - a.ph;
                    wi0[i] = wo * a.sw - a.pw;
                } else {
                    pix0[i] = 0;
                    hi0[i] = -(1 << 28);   // never in range
                    wi0[i] = 0;
                }
            }
            for (int kc = 0; kc < nk; ++kc) {
                const int k = kc * BLOCK_K + j * 8;
                const bool k_ok = k < a.K;
                int kh = 0, kw = 0, c = 0;
                if (k_ok) {
                    const int tap = k / a.Cin;
                    c = k - tap * a.Cin;
                    kh = tap / a.KW;
                    kw = tap - kh * a.KW;
                }
                // ---- issue all global loads of this stage first
                uint4 av[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int hi = hi0[i] + kh * a.dh, wi = wi0[i] + kw * a.dw;
                    if (k_ok && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) {
                        const long long off = ((long long)pix0[i] + (long long)hi * a.W + wi) * a.in_ld + a.in_choff + c;
                        av[i] = ldg16(x + off);
                    } else {
                        av[i] = make_uint4(0u, 0u, 0u, 0u);
                        if (a.pro_scale != nullptr) av[i].x = 0xFFFFFFFFu, av[i].y = 0x7FC07FC0u;   // tag: padding
                    }
                }
                uint4 bv[BLOCK_N / 16];
#pragma unroll
                for (int i = 0; i < BLOCK_N / 16; ++i) {
                    const int n = n0 + r0 + 16 * i;
                    bv[i] = (k_ok && n < a.Cout) ? ldg16(w + (long long)n * a.K + k) : make_uint4(0u, 0u, 0u, 0u);
                }
                // ---- optional prologue: relu(x*scale + shift) per input channel (zero padding stays 0)
                if (a.pro_scale != nullptr) {
                    float sc[8], sh[8];
                    if (k_ok) {
                        const float4 s0 = __ldg(reinterpret_cast<const float4 *>(a.pro_scale + c));
                        const float4 s1 = __ldg(reinterpret_cast<const float4 *>(a.pro_scale + c + 4));
                        const float4 h0 = __ldg(reinterpret_cast<const float4 *>(a.pro_shift + c));
                        const float4 h1 = __ldg(reinterpret_cast<const float4 *>(a.pro_shift + c + 4));
                        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
                        sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) sc[q] = sh[q] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (av[i].x == 0xFFFFFFFFu && av[i].y == 0x7FC07FC0u) {   // padded / out of range
                            av[i] = make_uint4(0u, 0u, 0u, 0u);
                            continue;
                        }
                        uint32_t *pv = reinterpret_cast<uint32_t *>(&av[i]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float2 f = unpack2(pv[q]);
                            f.x = fmaf(f.x, sc[2 * q], sh[2 * q]);
                            f.y = fmaf(f.y, sc[2 * q + 1], sh[2 * q + 1]);
                            if (a.pro_relu) {
                                f.x = fmaxf(f.x, 0.f);
                                f.y = fmaxf(f.y, 0.f);
                            }
                            pv[q] = pack2(f.x, f.y);
                        }
                    }
                }
                // ---- wait for the slot, then fill it
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sa = base + stage * C::kStageBytes + sw_off;
                const uint32_t sb = sa + C::kABytes;
#pragma unroll
                for (int i = 0; i < 8; ++i) sts16(sa + i * 2048, av[i]);
#pragma unroll
                for (int i = 0; i < BLOCK_N / 16; ++i) sts16(sb + i * 2048, bv[i]);
                fence_proxy_async();
                mbar_arrive(full_bar(stage));
                if (++stage == C::kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        constexpr uint32_t idesc = make_idesc(BLOCK_N);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, acc_phase = (it >> 1) & 1u;
            mbar_wait(acce_bar(buf), acc_phase ^ 1u);       // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
            for (int kc = 0; kc < nk; ++kc) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint64_t ad = make_desc_sw128(sa), bd = make_desc_sw128(sa + C::kABytes);
#pragma unroll
                    for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                        // advance 32 bytes (16 bf16) inside the swizzle atom: +2 in 16-byte units
                        umma_bf16(d_tmem, ad + 2u * kk, bd + 2u * kk, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));           // smem slot free once these MMAs retire
                    if (kc == nk - 1) umma_commit(accf_bar(buf));
                }
                __syncwarp();
                if (++stage == C::kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
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
            mbar_wait(accf_bar(buf), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + c0, r);
                tmem_ld_wait();
                const int n = n0 + c0;
                if (m < a.M && n < a.Cout) {
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
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] += to_f32(rp[e]);
                    }
                    apply_act_vec(v, a.act);
                    if (grow != nullptr) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4) {
                            const float4 g4 = __ldg(reinterpret_cast<const float4 *>(grow + n + e));
                            v[e] *= g4.x; v[e + 1] *= g4.y; v[e + 2] *= g4.z; v[e + 3] *= g4.w;
                        }
                    }
                    TOut *yp = y + m * a.out_ld + a.out_choff + n;
                    if constexpr (sizeof(TOut) == 2) {
                        uint4 o0 = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                        uint4 o1 = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
                        *reinterpret_cast<uint4 *>(yp) = o0;
                        *reinterpret_cast<uint4 *>(yp + 8) = o1;
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
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BLOCK_N, typename TOut, typename TRes>
int launch_one(const ConvArgs &a, cudaStream_t s) {
    using C = Cfg<BLOCK_N>;
    auto kern = conv_tc_kernel<BLOCK_N, TOut, TRes>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_tc) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    const long long mt = (a.M + BLOCK_M - 1) / BLOCK_M;
    const int ntn = (a.Cout + BLOCK_N - 1) / BLOCK_N;
    const long long tiles = mt * ntn;
    // co-resident CTAs share one SM's 512 TMEM columns and its shared memory
    int per_sm = 512 / C::kTmemCols;
    const int by_smem = (227 * 1024) / C::kSmemBytes;
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > tiles) grid = tiles;
    kern<<<(unsigned)grid, kThreads, C::kSmemBytes, s>>>(a, ntn, tiles);
    return check_launch("conv_tc_kernel");
}

template <typename TOut, typename TRes>
int launch_n(const ConvArgs &a, cudaStream_t s) {
    if (a.Cout <= 32) return launch_one<32, TOut, TRes>(a, s);
    if (a.Cout <= 64) return launch_one<64, TOut, TRes>(a, s);
    if (a.Cout <= 128) return launch_one<128, TOut, TRes>(a, s);
    return launch_one<256, TOut, TRes>(a, s);
}

}  // namespace

bool conv_tc_supported(const ConvArgs &a, int in_dtype) {
    if (in_dtype != SPK_DT_BF16) return false;
    if (a.Cin % 8 != 0 || a.in_ld % 8 != 0 || a.in_choff % 8 != 0) return false;
    if (a.Cout % 16 != 0 || a.out_ld % 8 != 0 || a.out_choff % 8 != 0) return false;
    if (a.res != nullptr && (a.res_ld % 8 != 0 || a.res_choff % 8 != 0)) return false;
    if (a.M < BLOCK_M) return false;          // tiny problems (e.g. per-segment dense) stay on CUDA cores
    return true;
}

int launch_conv_tc(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s) {
    if (a.M == 0) return SPK_OK;
    // the first-generation kernel below stays selectable for A/B runs (SPK_CONV_TC_V1=1)
    static const bool v1 = [] { const char *e = getenv("SPK_CONV_TC_V1"); return e && e[0] == '1'; }();
    if (!v1 || a.post_scale != nullptr || a.pad_reflect || a.gate_additive) return launch_conv_tc2(a, out_dtype, res_dtype, s);
    const bool res_bf16 = a.res == nullptr ? (out_dtype == SPK_DT_BF16) : (res_dtype == SPK_DT_BF16);
    if (out_dtype == SPK_DT_BF16) {
        if (!res_bf16) {
            set_error("conv_tc: a bf16 output takes a bf16 residual");
            return SPK_ERR_UNSUPPORTED;
        }
        return launch_n<bf16, bf16>(a, s);
    }
    return res_bf16 ? launch_n<float, bf16>(a, s) : launch_n<float, float>(a, s);
}

}  // namespace spk
