// Fused CAM layer for sm_100a: the dilated k=3 "local" conv of CAMLayer, the context gate and the
// gating multiply in ONE kernel (speakerlab/models/campplus/layers.py:93-110):
//
//   y   = linear_local(x)                         Conv1d(128 -> 32, k=3, dilation d, padding d)
//   ctx = mean_T(x) + seg_avg_100(x)              full-utterance mean + 100-frame segment means
//   m   = sigmoid(W2 relu(W1 ctx + b1) + b2)      1x1 convs with bias, evaluated once per window
//   out = y * m                                   written into the block's concat buffer slice
//
// A CTA item is a group of whole segments: their rows (zero padded by d on both sides, pitch P)
// are staged ONCE in shared memory as sixteen 16-byte channel planes by cp.async producer warps.
// From that one copy
//   * the MMA warp runs the three taps as shifted no-swizzle K-major UMMA descriptors
//     (tap k = start address + k*d rows), tcgen05.mma M=128 N=32 K=16, accumulators in TMEM;
//   * the epilogue warps - idle until the accumulator is ready - reduce the column sums of the
//     same staged rows (fixed order: deterministic), run the two tiny mat-vecs of the gate, and
//     then multiply the gate into the accumulator on its way to HBM.
// Compared with the unfused path this removes one kernel launch, the separate read of x for the
// context and the 3x gather of x for the taps.  Slabs and TMEM accumulators are double buffered,
// so staging, MMA and epilogue of consecutive items overlap.
#include <algorithm>
#include <mutex>

#include "ops.cuh"

namespace spk {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kCin = 128, kCout = 32, kPlanes = kCin / 8, kTaps = 3;
constexpr int kProd = 128, kEpi = 256, kThreads = kProd + 32 + kEpi;       // 416
constexpr int kMaxSeg = 8, kMaxWin = 4, kMaxHidden = 64;
constexpr uint32_t kSpinLimit = 1u << 26;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
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
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
constexpr uint32_t kIdescN32 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
__device__ __forceinline__ uint4 ldg16(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ float2 unpack2(uint32_t v) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162 *>(&v);
    return make_float2(__low2float(t), __high2float(t));
}

struct CamGeom {
    int T, d, P, G, n_tiles, px, nwin, seg_len, hidden;
    int smem_bytes, tmem_cols;
    unsigned p_magic;        // ceil(2^32 / P)
    uint32_t off_w, off_slab, slab_bytes, off_part, off_win, off_tot, off_hid, off_gate, off_bar;
};

__global__ void __launch_bounds__(kThreads, 1)
cam_local_kernel(const ConvArgs a, const CamGeom g, const float *__restrict__ w1t, const float *__restrict__ b1,
                 const float *__restrict__ w2t, const float *__restrict__ b2, int n_items) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t s0 = smem_u32(smem);
    const uint32_t s_w = s0 + g.off_w, s_slab0 = s0 + g.off_slab, s_bar = s0 + g.off_bar;
    float *part = reinterpret_cast<float *>(smem + g.off_part);     // [2][16][128]
    float *win = reinterpret_cast<float *>(smem + g.off_win);       // [2][G][nwin][128]   (per slab buffer)
    float *hid = reinterpret_cast<float *>(smem + g.off_hid);       // [G*nwin][hidden]
    float *gate = reinterpret_cast<float *>(smem + g.off_gate);     // [2][G][nwin][32]
    auto sfull = [&](int i) { return s_bar + 8u * i; };
    auto sempty = [&](int i) { return s_bar + 8u * (2 + i); };
    auto afull = [&](int i) { return s_bar + 8u * (4 + i); };
    auto aempty = [&](int i) { return s_bar + 8u * (6 + i); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + g.off_bar + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t plane = (uint32_t)g.px * 16u;
    const uint32_t acc_cols = (uint32_t)g.n_tiles * 32u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(sfull(i), kProd);
            mbar_init(sempty(i), 1 + kEpi);        // MMA commit + every epilogue thread (they read the slab for the context)
            mbar_init(afull(i), 1);
            mbar_init(aempty(i), kEpi);
        }
        fence_barrier_init();
    }
    if (warp == 4) {
        __syncwarp();
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), g.tmem_cols);
    }
    {   // weights -> smem as [tap][chunk c][n] x 16 B; zero the slack rows of both slabs once
        const bf16 *w = static_cast<const bf16 *>(a.w);          // [32][3][128]
        for (int idx = threadIdx.x; idx < kCout * kTaps * kPlanes; idx += kThreads) {
            const int c = idx % kPlanes, t = (idx / kPlanes) % kTaps, n = idx / (kPlanes * kTaps);
            sts16(s_w + (uint32_t)(((t * kPlanes + c) * 32 + n) * 16), ldg16(w + ((long long)n * kTaps + t) * kCin + c * 8));
        }
        const int rows = g.G * g.P, slack = g.px - rows;
        for (int idx = threadIdx.x; idx < 2 * slack * kPlanes; idx += kThreads) {
            const int c = idx % kPlanes;
            int p = idx / kPlanes;
            const uint32_t sb = s_slab0 + (p >= slack ? g.slab_bytes : 0u);
            if (p >= slack) p -= slack;
            sts16(sb + c * plane + (uint32_t)(rows + p) * 16u, make_uint4(0u, 0u, 0u, 0u));
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =========================== producers (cp.async, zero-fill for the padding rows) ===========================
        const bf16 *x = static_cast<const bf16 *>(a.x);
        const int pieces = g.G * g.P * kPlanes;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b0 = item * g.G;
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(sempty(buf), ph ^ 1u);
            const uint32_t sb = s_slab0 + buf * g.slab_bytes;
            for (int idx = threadIdx.x; idx < pieces; idx += kProd) {
                const int c = idx & (kPlanes - 1);
                const int row = idx >> 4;
                const int gs = (int)__umulhi((unsigned)row, g.p_magic);
                const int u = row - gs * g.P;
                const int t = u - g.d;
                const bool ok = (b0 + gs < a.B) && t >= 0 && t < g.T;
                const bf16 *src = ok ? x + ((long long)(b0 + gs) * g.T + t) * a.in_ld + a.in_choff + c * 8 : x;
                cp_async16(sb + c * plane + (uint32_t)row * 16u, src, ok ? 16u : 0u);
            }
            cp_async_wait_all();
            fence_proxy_async();
            mbar_arrive(sfull(buf));
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(aempty(buf), ph ^ 1u);
            mbar_wait(sfull(buf), ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sb = s_slab0 + buf * g.slab_bytes;
                for (int t = 0; t < g.n_tiles; ++t) {
                    const uint32_t dcol = tmem_base + buf * acc_cols + (uint32_t)t * 32u;
#pragma unroll
                    for (int k = 0; k < kTaps; ++k) {
                        const uint32_t a_row = sb + (uint32_t)(t * 128 + k * g.d) * 16u;
                        const uint32_t b_tap = s_w + (uint32_t)(k * kPlanes * 32 * 16);
#pragma unroll
                        for (int j = 0; j < kCin / 16; ++j)
                            umma_bf16(dcol, make_desc_nosw(a_row + 2u * j * plane, plane, 128u),
                                      make_desc_nosw(b_tap + 2u * j * 512u, 512u, 128u), kIdescN32, (k | j) ? 1u : 0u);
                    }
                }
                umma_commit(sempty(buf));
                umma_commit(afull(buf));
            }
            __syncwarp();
        }
    } else {
        // =========================== context gate + epilogue (256 threads) ===========================
        const int et = threadIdx.x - (kProd + 32);          // 0..255
        const int ew = et >> 5;                              // 0..7
        const int q = warp & 3, tsel = ew >> 2;
        bf16 *y = static_cast<bf16 *>(a.y);
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b0 = item * g.G;
            const int g_valid = min(g.G, a.B - b0);
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            const uint32_t sb = s_slab0 + buf * g.slab_bytes;
            float *winb = win + (size_t)buf * g.G * g.nwin * kCin;
            float *gateb = gate + (size_t)buf * g.G * g.nwin * kCout;
            mbar_wait(sfull(buf), ph);
            const int combos = g_valid * g.nwin;
            // ---- column sums of the staged rows: thread (plane c = et & 15, row group rg = et >> 4) sums its
            // rows of one (segment, window), the 16 row groups are folded in a fixed order (deterministic).
            // `part` is double buffered so each round needs one barrier only.
            {
                const int c = et & 15, rg = et >> 4;
                for (int cb = 0; cb < combos; ++cb) {
                    const int gs = cb / g.nwin, w = cb - gs * g.nwin;
                    const int t0 = w * g.seg_len, t1 = min(g.T, t0 + g.seg_len);
                    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int t = t0 + rg; t < t1; t += 16) {
                        const uint4 v = lds16(sb + c * plane + (uint32_t)(gs * g.P + g.d + t) * 16u);
                        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 f = unpack2(vv[h]);
                            acc[2 * h] += f.x;
                            acc[2 * h + 1] += f.y;
                        }
                    }
                    float *pp = part + (cb & 1) * 16 * kCin;
                    *reinterpret_cast<float4 *>(&pp[rg * kCin + c * 8]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    *reinterpret_cast<float4 *>(&pp[rg * kCin + c * 8 + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                    epi_bar_sync();
                    if (et < kCin) {
                        float s = 0.f;
#pragma unroll
                        for (int r = 0; r < 16; ++r) s += pp[r * kCin + et];
                        winb[cb * kCin + et] = s;
                    }
                }
            }
            // the slab is no longer needed by these threads
            mbar_arrive(sempty(buf));
            // ---- ctx = tot/T + win/len, in place (thread et owns channel et of every combo: no barrier needed)
            if (et < kCin) {
                for (int gs = 0; gs < g_valid; ++gs) {
                    float t = 0.f;
                    for (int w = 0; w < g.nwin; ++w) t += winb[(gs * g.nwin + w) * kCin + et];
                    t /= (float)g.T;
                    for (int w = 0; w < g.nwin; ++w) {
                        const int len = min(g.T, (w + 1) * g.seg_len) - w * g.seg_len;
                        float *p = &winb[(gs * g.nwin + w) * kCin + et];
                        *p = t + *p / (float)len;
                    }
                }
            }
            epi_bar_sync();
            // ---- hidden = relu(W1 ctx + b1): thread (j = et & 63, quarter qd = et >> 6) does 32 channels of
            // every combo from the TRANSPOSED weights (coalesced rows of 64 floats); quarters folded in order
            {
                const int j = et & 63, qd = et >> 6;
                float acc[kMaxSeg * kMaxWin > 8 ? 8 : kMaxSeg * kMaxWin];
                float *hp = part;                            // [4][combos][64] aliases the (now dead) row-group partials
                if (j < g.hidden) {
#pragma unroll
                    for (int cb = 0; cb < 8; ++cb) acc[cb] = 0.f;
                    for (int i = 0; i < 32; ++i) {
                        const int ch = qd * 32 + i;
                        const float wv = __ldg(w1t + (size_t)ch * g.hidden + j);
#pragma unroll
                        for (int cb = 0; cb < 8; ++cb)
                            if (cb < combos) acc[cb] = fmaf(wv, winb[cb * kCin + ch], acc[cb]);
                    }
#pragma unroll
                    for (int cb = 0; cb < 8; ++cb)
                        if (cb < combos) hp[(qd * 8 + cb) * kMaxHidden + j] = acc[cb];
                }
                epi_bar_sync();
                for (int idx = et; idx < combos * g.hidden; idx += kEpi) {
                    const int cb = idx / g.hidden, jj = idx - cb * g.hidden;
                    const float s = ((hp[(0 * 8 + cb) * kMaxHidden + jj] + hp[(1 * 8 + cb) * kMaxHidden + jj]) +
                                     (hp[(2 * 8 + cb) * kMaxHidden + jj] + hp[(3 * 8 + cb) * kMaxHidden + jj])) + __ldg(b1 + jj);
                    hid[cb * kMaxHidden + jj] = fmaxf(s, 0.f);
                }
                epi_bar_sync();
            }
            // ---- gate = sigmoid(W2 hidden + b2): thread (o = et & 31, part pt = et >> 5) does 8 hidden units
            {
                const int o = et & 31, pt = et >> 5;
                float *gp = part;                            // [8][combos][32]
                float acc[8];
#pragma unroll
                for (int cb = 0; cb < 8; ++cb) acc[cb] = 0.f;
                for (int i = 0; i < 8; ++i) {
                    const int jj = pt * 8 + i;
                    if (jj < g.hidden) {
                        const float wv = __ldg(w2t + (size_t)jj * kCout + o);
#pragma unroll
                        for (int cb = 0; cb < 8; ++cb)
                            if (cb < combos) acc[cb] = fmaf(wv, hid[cb * kMaxHidden + jj], acc[cb]);
                    }
                }
#pragma unroll
                for (int cb = 0; cb < 8; ++cb)
                    if (cb < combos) gp[(pt * 8 + cb) * kCout + o] = acc[cb];
                epi_bar_sync();
                for (int idx = et; idx < combos * kCout; idx += kEpi) {
                    const int cb = idx >> 5, oo = idx & 31;
                    float s = __ldg(b2 + oo);
#pragma unroll
                    for (int q8 = 0; q8 < 8; ++q8) s += gp[(q8 * 8 + cb) * kCout + oo];
                    gateb[cb * kCout + oo] = 1.f / (1.f + expf(-s));
                }
                epi_bar_sync();
            }
            // ---- accumulator -> gate -> bf16 -> concat buffer slice
            mbar_wait(afull(buf), ph);
            tc_fence_after();
            for (int t = tsel; t < g.n_tiles; t += 2) {
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
#pragma unroll
                    for (int e = 0; e < 32; e += 8)
                        *reinterpret_cast<uint4 *>(yp + e) =
                            make_uint4(pack2(v[e], v[e + 1]), pack2(v[e + 2], v[e + 3]), pack2(v[e + 4], v[e + 5]), pack2(v[e + 6], v[e + 7]));
                }
            }
            tc_fence_before();
            mbar_arrive(aempty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, g.tmem_cols);
}

bool geometry(const ConvArgs &a, int hidden, int seg_len, CamGeom &g) {
    g.T = a.W; g.d = a.dw; g.seg_len = seg_len; g.hidden = hidden;
    g.nwin = (g.T + seg_len - 1) / seg_len;
    g.P = (g.T + 2 * g.d + 7) & ~7;
    g.G = std::min(kMaxSeg, 248 / g.P);
    if (g.G < 1 || g.nwin > kMaxWin || hidden > kMaxHidden || hidden < 1) return false;
    while (g.G * g.nwin > 8) --g.G;             // the gate MLP keeps at most 8 (segment, window) combos in registers
    g.n_tiles = (g.G * g.P + 127) / 128;
    // rows the shifted views can touch: n_tiles*128 + 2d; planes padded to px = 1 (mod 8) rows so the
    // sixteen planes of one row fall into different banks for the cp.async stores
    int px = std::max(g.G * g.P, g.n_tiles * 128) + 2 * g.d + 8;
    px = ((px + 7) & ~7) + 1;
    g.px = px;
    g.p_magic = (unsigned)(((1ull << 32) + g.P - 1) / g.P);
    g.slab_bytes = (uint32_t)px * 16u * kPlanes;
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off += (bytes + 127u) & ~127u; return o; };
    g.off_w = take(kTaps * kPlanes * 32 * 16);
    g.off_slab = take(2 * g.slab_bytes);
    g.off_part = take(2 * 16 * kCin * 4);       // double-buffered row-group partials; later aliased by the MLP partials
    g.off_win = take(2 * g.G * g.nwin * kCin * 4);
    g.off_tot = take(g.G * kCin * 4);
    g.off_hid = take(g.G * g.nwin * kMaxHidden * 4);
    g.off_gate = take(2 * g.G * g.nwin * kCout * 4);
    g.off_bar = take(128);
    g.smem_bytes = (int)off;
    const int cols = 2 * g.n_tiles * 32;
    g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    return g.smem_bytes <= 220 * 1024 && cols <= 512;
}

}  // namespace

bool cam_local_supported(const ConvArgs &a, int in_dtype, int out_dtype, int hidden, int seg_len) {
    if (in_dtype != SPK_DT_BF16 || out_dtype != SPK_DT_BF16) return false;
    if (a.Cin != kCin || a.Cout != kCout || a.KH != 1 || a.KW != kTaps || a.H != 1 || a.Ho != 1) return false;
    if (a.sw != 1 || a.sh != 1 || a.pw != a.dw || a.dw < 1 || a.Wo != a.W) return false;
    if (a.pro_scale != nullptr || a.epi_scale != nullptr || a.res != nullptr || a.act != SPK_ACT_NONE) return false;
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
        attr_err = cudaFuncSetAttribute(cam_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(cam_local) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    const int items = (a.B + g.G - 1) / g.G;
    const int grid = std::min(items, sm_count());
    cam_local_kernel<<<grid, kThreads, g.smem_bytes, s>>>(a, g, w1t, b1, w2t, b2, items);
    return check_launch("cam_local_kernel");
}

}  // namespace spk
