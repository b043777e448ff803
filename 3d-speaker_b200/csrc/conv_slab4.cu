// Slab kernel for 3x3 stride-1 convs with up to 64 channels: the Res2Net branch convs of ERes2NetV2 / ERes2Net
// (speakerlab/models/eres2net/ERes2NetV2.py:65-91: `convs[i]` on `width` = 26/52 resp. 24/48 channels, padded to 32/64/48)
// and the 32-channel FCM convs of CAM++ on segments wider than 254 frames, which conv_slab3 does not take.
//
// Same idea as conv_slab3.cu - the input band of an item is staged ONCE by a TMA box, every tap of the conv is that slab
// seen through a pixel-shifted UMMA descriptor, the output band leaves through a swizzled staging buffer and one TMA box
// store - generalised in three directions:
//   * a staged pixel is a 64-byte (<= 32 channels, 64B swizzle) or 128-byte (<= 64 channels, 128B swizzle) row; the
//     MMA is M=128, N=Cout, K=16 with Cin/16 K steps per tap.  At N=64 one MMA reads 6 KB of shared memory for 32 cycles
//     of tensor work (N=32: 5 KB for 16), so these convs can run at up to 2/3 of the tensor peak;
//   * an item is (segment, band of R rows, column part): rows wider than the 256-pixel TMA box - or too wide for a
//     useful band to fit shared memory next to the 72 KB of weights - are cut into parts with a one-column halo; each
//     part stores through its own tensor map whose width ends at the part's last column, so the halo columns of the
//     staged band are clipped instead of overwriting the neighbour part;
//   * channel windows: the input box may cover more channels than Cin (never multiplied: only Cin/16 K steps are
//     issued); the output map ends at the last output channel, so a 64-channel box into a 48-channel window is clipped.
// The generic gather kernel these convs used before ran the tensor pipe at 2 % (profiles/r01_ncu_full_tc2.md).
//
//   warp 0      TMA producer        warp 1      MMA issuer (+ TMEM allocation)
//   warps 2-9   epilogue (two warps per TMEM lane quarter, alternating tiles)      warp 10     TMA store
#include <algorithm>
#include <cmath>
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
constexpr int kEpi = 256, kThreads = 64 + kEpi + 32;  // 352
constexpr int kMaxParts = 8;

struct Slab4Geom {
    int Wp, Wt, P, R, n_tiles, n_bands, rows, px, nbuf, nstg, tmem_cols, N, ksteps;
    uint32_t off_w, off_ss, off_stg, stg_bytes, off_slab, slab_bytes, off_bar;
    int smem_bytes;
};

struct Slab4Maps {
    CUtensorMap x, r, y[kMaxParts];
};

template <int CP> struct Pix {
    static constexpr uint32_t kBytes = CP * 2;                  // bytes per staged pixel
    static constexpr uint32_t kUnits = kBytes / 16;             // 16-byte descriptor units per pixel
    static constexpr uint32_t kLayout = CP == 32 ? kLayoutSw64 : kLayoutSw128;
    static constexpr uint32_t kSbo = 8 * kBytes;                // 8-pixel groups
    static constexpr int kSwizzle = CP == 32 ? 64 : 128;
    // position of 16-byte chunk c of row (pixel) p
    static __device__ __forceinline__ uint32_t chunk(uint32_t c, uint32_t p) { return CP == 32 ? (c ^ ((p >> 1) & 3u)) : (c ^ (p & 7u)); }
};

template <int CP>
__global__ void __launch_bounds__(kThreads, 1)
conv_slab4_kernel(const ConvArgs a, const Slab4Geom g, long long n_items, const __grid_constant__ Slab4Maps maps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using PX = Pix<CP>;
    constexpr int KSMAX = CP / 16;
    const uint32_t s0 = smem_u32(smem);
    const uint32_t s_w = s0 + g.off_w, s_ss = s0 + g.off_ss, s_slab0 = s0 + g.off_slab, s_bar = s0 + g.off_bar, s_stg = s0 + g.off_stg;
    auto sfull = [&](uint32_t i) { return s_bar + 8u * i; };
    auto sempty = [&](uint32_t i) { return s_bar + 8u * (4 + i); };
    auto afull = [&](uint32_t i) { return s_bar + 8u * (8 + i); };
    auto aempty = [&](uint32_t i) { return s_bar + 8u * (10 + i); };
    auto rfull = [&](uint32_t i) { return s_bar + 8u * (12 + i); };      // residual band landed in the staging buffer
    auto gfull = [&](uint32_t i) { return s_bar + 8u * (14 + i); };      // staging buffer holds the finished band
    auto gfree = [&](uint32_t i) { return s_bar + 8u * (16 + i); };      // the TMA store has read the staging buffer
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + g.off_bar + 160);
    const bool has_res = a.res != nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t acc_cols = (uint32_t)(g.n_tiles * g.N);
    const int per_seg = g.n_bands * g.P;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tmap_prefetch(&maps.x);
        for (uint32_t i = 0; i < 4; ++i) {
            mbar_init(sfull(i), 1);
            mbar_init(sempty(i), 1);
        }
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(afull(i), 1);
            mbar_init(aempty(i), kEpi);
            mbar_init(rfull(i), 1);
            mbar_init(gfull(i), kEpi);
            mbar_init(gfree(i), 1);
        }
        for (int i = 0; i < g.P; ++i) tmap_prefetch(&maps.y[i]);
        if (has_res) tmap_prefetch(&maps.r);
        fence_barrier_init();
    }
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_slot)), g.tmem_cols);
    }
    {   // weights -> smem: per tap a K-major swizzled [N cout][CP cin] block; epilogue scale/shift; zero the pixels
        // behind the last staged row of each slab (TMA never writes them, the last tile's shifted views read them)
        const bf16 *w = static_cast<const bf16 *>(a.w);      // [Cout][9][Cin]
        const int cin8 = a.Cin / 8;
        for (int idx = threadIdx.x; idx < g.N * 9 * cin8; idx += kThreads) {
            const int c = idx % cin8, t = (idx / cin8) % 9, n = idx / (9 * cin8);
            sts16(s_w + (uint32_t)t * (uint32_t)g.N * PX::kBytes + (uint32_t)n * PX::kBytes + (PX::chunk((uint32_t)c, (uint32_t)n) << 4),
                  ldg16(w + ((long long)n * 9 + t) * a.Cin + c * 8));
        }
        float *ss = reinterpret_cast<float *>(smem + g.off_ss);
        for (int n = threadIdx.x; n < g.N; n += kThreads) {
            ss[n] = a.epi_scale != nullptr ? __ldg(a.epi_scale + n) : 1.f;
            ss[64 + n] = a.epi_shift != nullptr ? __ldg(a.epi_shift + n) : 0.f;
        }
        const int slack = g.px - g.rows * g.Wp;
        for (int idx = threadIdx.x; idx < g.nbuf * slack * (int)PX::kUnits; idx += kThreads) {
            const int c = idx % (int)PX::kUnits;
            int p = idx / (int)PX::kUnits;
            const int bi = p / slack;
            p -= bi * slack;
            sts16(s_slab0 + (uint32_t)bi * g.slab_bytes + (uint32_t)(g.rows * g.Wp + p) * PX::kBytes + (uint32_t)c * 16u, make_uint4(0u, 0u, 0u, 0u));
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // everything above read only parameters; the activations come from the previous kernel

    auto decode = [&](long long item, int &b, int &h0, int &w0, int &part) {
        b = (int)(item / per_seg);
        const int r = (int)(item - (long long)b * per_seg);
        const int band = r / g.P;
        part = r - band * g.P;
        h0 = band * g.R;
        w0 = part * g.Wt;
    };

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (elect_one()) {
            const uint32_t bytes = PX::kBytes * (uint32_t)(g.rows * g.Wp);
            uint32_t buf = 0, ph = 0, it = 0;
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                int b, h0, w0, part;
                decode(item, b, h0, w0, part);
                mbar_wait(sempty(buf), ph ^ 1u);
                mbar_arrive_expect_tx(sfull(buf), bytes);
                tmap_load_4d(s_slab0 + buf * g.slab_bytes, &maps.x, a.in_choff, w0 - 1, h0 - 1, b, sfull(buf));
                if (has_res) {      // residual band -> staging buffer, once the store that last used it has read it
                    const uint32_t gb = it % (uint32_t)g.nstg, gph = (it / (uint32_t)g.nstg) & 1u;
                    mbar_wait(gfree(gb), gph ^ 1u);
                    mbar_arrive_expect_tx(rfull(gb), PX::kBytes * (uint32_t)(g.R * g.Wp));
                    tmap_load_4d(s_stg + gb * g.stg_bytes, &maps.r, a.res_choff, w0, h0, b, rfull(gb));
                }
                if (++buf == (uint32_t)g.nbuf) { buf = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        const uint32_t hi = desc_hi(PX::kSbo, PX::kLayout);
        const uint32_t wpu = (uint32_t)g.Wp * PX::kUnits;               // one slab row, in 16-byte descriptor units
        const uint32_t tapu = (uint32_t)g.N * PX::kUnits;               // one weight tap
        const uint32_t idesc = idesc_bf16(g.N);
        const int ksteps = g.ksteps;
        uint32_t it = 0, sbuf = 0, sph = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(aempty(buf), ph ^ 1u);
            mbar_wait(sfull(sbuf), sph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t lo_s = desc_lo(s_slab0 + sbuf * g.slab_bytes, 16u);
                const uint32_t lo_w = desc_lo(s_w, 16u);
                uint32_t d = tmem_base + buf * acc_cols;
                uint32_t tile = 0;                                // 128 pixels
                for (int t = 0; t < g.n_tiles; ++t, d += (uint32_t)g.N, tile += 128u * PX::kUnits) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const uint32_t lo_a = lo_s + tile + kh * wpu + (uint32_t)kw * PX::kUnits;
                            const uint32_t lo_b = lo_w + (uint32_t)(kh * 3 + kw) * tapu;
#pragma unroll
                            for (int ks = 0; ks < KSMAX; ++ks) {
                                if (ks < ksteps) {
                                    if (kh == 0 && kw == 0 && ks == 0) umma_bf16(d, desc64(lo_a, hi), desc64(lo_b, hi), idesc, 0u);
                                    else umma_bf16_acc(d, desc64(lo_a + 2u * ks, hi), desc64(lo_b + 2u * ks, hi), idesc);
                                }
                            }
                        }
                }
                umma_commit(sempty(sbuf));     // slab reusable once these MMAs retire
                umma_commit(afull(buf));
            }
            __syncwarp();
            if (++sbuf == (uint32_t)g.nbuf) { sbuf = 0; sph ^= 1u; }
        }
    } else if (warp < 10) {
        // =========================== epilogue ===========================
        const int q = warp & 3;
        const int tsel = (warp - 2) >> 2;           // 0 or 1: even / odd tiles
        const int nsteps = g.N / 16;
        uint32_t it = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            const uint32_t gb = it % (uint32_t)g.nstg, gph = (it / (uint32_t)g.nstg) & 1u;
            const uint32_t stg = s_stg + gb * g.stg_bytes;
            // the staging buffer is ours once the residual has landed in it (which implies the previous store has read
            // it) or, without a residual, once that store has read it
            if (has_res) mbar_wait(rfull(gb), gph);
            else mbar_wait(gfree(gb), gph ^ 1u);
            mbar_wait(afull(buf), ph);
            tc_fence_after();
            for (int t = tsel; t < g.n_tiles; t += 2) {
                const uint32_t p = (uint32_t)(t * 128 + q * 32 + lane);               // slab pixel == staging pixel
                const uint32_t taddr = tmem_base + buf * acc_cols + (uint32_t)(t * g.N) + ((uint32_t)(q * 32) << 16);
                const uint32_t prow = stg + p * PX::kBytes;
                uint32_t r[KSMAX][16];
#pragma unroll
                for (int s = 0; s < KSMAX; ++s)
                    if (s < nsteps) tmem_ld16(taddr + 16u * s, r[s]);
                tmem_ld_wait();
#pragma unroll
                for (int s = 0; s < KSMAX; ++s) {
                    if (s >= nsteps) break;
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const uint4 sc = lds16(s_ss + (uint32_t)(16 * s + e) * 4u), sh = lds16(s_ss + 256u + (uint32_t)(16 * s + e) * 4u);
                        v[e] = fmaf(__uint_as_float(r[s][e]), __uint_as_float(sc.x), __uint_as_float(sh.x));
                        v[e + 1] = fmaf(__uint_as_float(r[s][e + 1]), __uint_as_float(sc.y), __uint_as_float(sh.y));
                        v[e + 2] = fmaf(__uint_as_float(r[s][e + 2]), __uint_as_float(sc.z), __uint_as_float(sh.z));
                        v[e + 3] = fmaf(__uint_as_float(r[s][e + 3]), __uint_as_float(sc.w), __uint_as_float(sh.w));
                    }
                    const uint32_t a0 = prow + (PX::chunk(2u * s, p) << 4), a1 = prow + (PX::chunk(2u * s + 1u, p) << 4);
                    if (has_res) {
                        const uint4 w0 = lds16(a0), w1 = lds16(a1);
                        const uint32_t w8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            const float2 f = unpack2(w8[h]);
                            v[2 * h] += f.x;
                            v[2 * h + 1] += f.y;
                        }
                    }
                    apply_act_vec(v, a.act);
                    sts16(a0, make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7])));
                    sts16(a1, make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15])));
                }
            }
            tc_fence_before();
            mbar_arrive(aempty(buf));
            fence_proxy_async();            // staging writes (generic proxy) -> TMA store (async proxy)
            mbar_arrive(gfull(gb));
        }
    } else {
        // =========================== TMA store ===========================
        if (elect_one()) {
            uint32_t it = 0;
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                int b, h0, w0, part;
                decode(item, b, h0, w0, part);
                const uint32_t gb = it % (uint32_t)g.nstg, gph = (it / (uint32_t)g.nstg) & 1u;
                mbar_wait(gfull(gb), gph);
                tmap_store_4d(&maps.y[part], a.out_choff, w0, h0, b, s_stg + gb * g.stg_bytes);
                bulk_commit();
                bulk_wait_read0();          // the box has been read out of shared memory
                mbar_arrive(gfree(gb));
            }
            bulk_wait_all();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, g.tmem_cols);
}

int chan_pad(const ConvArgs &a) { return std::max(a.Cin, a.Cout) <= 32 ? 32 : 64; }

bool geometry(const ConvArgs &a, Slab4Geom &g) {
    const int CP = chan_pad(a), PB = CP * 2;
    const int N = a.Cout;
    double best = -1.0;
    const int budget = 227 * 1024 - 256;
    for (int P = 1; P <= kMaxParts; ++P) {
        const int Wt = (a.W + P - 1) / P;
        if (P > 1 && Wt < 16) break;
        const int Wp = (Wt + 2 + 7) & ~7;
        if (Wp > 256) continue;
        for (int R = std::min(a.Ho, 32); R >= 1; --R) {
            const int n_tiles = (R * Wp + 127) / 128;
            if (2 * n_tiles * N > 512 || R + 2 > 256) continue;
            const int rows = R + 2;
            const int px = (std::max(rows * Wp, n_tiles * 128 + 2 * Wp + 2) + 7) & ~7;
            const uint32_t slab_bytes = ((uint32_t)(px * PB) + 1023u) & ~1023u;
            const uint32_t stg_bytes = (uint32_t)(n_tiles * 128 * PB);
            const uint32_t wbytes = ((uint32_t)(9 * N * PB) + 1023u) & ~1023u;
            const uint32_t fixed = wbytes + 1024u;          // + scale/shift
            int nstg = 2;
            if (fixed + 2 * slab_bytes + 2 * stg_bytes > (uint32_t)budget) nstg = 1;
            if (fixed + 2 * slab_bytes + (uint32_t)nstg * stg_bytes > (uint32_t)budget) continue;
            // useful output pixels per tile pixel, with a mild preference for less halo re-reading and for two staging buffers
            const double eff = (double)R * ((double)a.W / P) / (n_tiles * 128.0) * std::pow((double)R / (R + 2), 0.3) * (nstg == 2 ? 1.0 : 0.93);
            if (eff > best + 1e-9) {
                best = eff;
                g.Wp = Wp; g.Wt = Wt; g.P = P; g.R = R; g.n_tiles = n_tiles; g.rows = rows; g.px = px; g.nstg = nstg;
                g.slab_bytes = slab_bytes; g.stg_bytes = stg_bytes;
                g.off_w = 0; g.off_ss = wbytes; g.off_stg = wbytes + 1024u; g.off_slab = g.off_stg + (uint32_t)nstg * stg_bytes;
            }
        }
    }
    if (best < 0) return false;
    g.N = N;
    g.ksteps = a.Cin / 16;
    g.nbuf = 2;
    while (g.nbuf < 4 && g.off_slab + (uint32_t)(g.nbuf + 1) * g.slab_bytes + 256u <= 227u * 1024u) ++g.nbuf;
    g.off_bar = g.off_slab + (uint32_t)g.nbuf * g.slab_bytes;
    g.smem_bytes = (int)g.off_bar + 256;
    g.n_bands = (a.Ho + g.R - 1) / g.R;
    const int cols = g.n_tiles * N * 2;
    g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    return g.smem_bytes <= 227 * 1024;
}

// {CP channels, Wp pixels, rows, 1 segment} boxes over a [B][H][W][ld] bf16 buffer whose channel axis ends at ch_end and
// whose width ends at w_end (everything beyond reads as zero / is not written)
int make_map(const void *ptr, int ld, int ch_end, int w_end, int W, int H, int B, int cp, int wp, int rows, CUtensorMap *out) {
    typedef std::tuple<const void *, int, int, int, int, int, int, int, int, int> Key;
    static std::mutex mu;
    static std::map<Key, CUtensorMap> cache;
    const Key key(ptr, ld, ch_end, w_end, W, H, B, cp, wp, rows);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return SPK_OK;
    }
    const uint64_t dims[4] = {(uint64_t)ch_end, (uint64_t)w_end, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
    const uint32_t box[4] = {(uint32_t)cp, (uint32_t)wp, (uint32_t)rows, 1u};
    const int rc = tmap_encode_bf16(ptr, 4, dims, strides, box, cp == 32 ? 64 : 128, out);
    if (rc != SPK_OK) return rc;
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return SPK_OK;
}

template <int CP>
int launch(const ConvArgs &a, const Slab4Geom &g, cudaStream_t s) {
    auto kern = conv_slab4_kernel<CP>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    if (attr_err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(conv_slab4) failed: %s", cudaGetErrorString(attr_err));
        return SPK_ERR_CUDA;
    }
    Slab4Maps maps;
    int rc = make_map(a.x, a.in_ld, std::min(a.in_ld, a.in_choff + CP), a.W, a.W, a.H, a.B, CP, g.Wp, g.rows, &maps.x);
    if (rc != SPK_OK) return rc;
    for (int p = 0; p < kMaxParts; ++p) {
        const int w_end = std::min(a.Wo, (std::min(p, g.P - 1) + 1) * g.Wt);
        rc = make_map(a.y, a.out_ld, a.out_choff + a.Cout, w_end, a.Wo, a.Ho, a.B, CP, g.Wp, g.R, &maps.y[p]);
        if (rc != SPK_OK) return rc;
    }
    maps.r = maps.y[0];
    if (a.res != nullptr) {
        rc = make_map(a.res, a.res_ld, std::min(a.res_ld, a.res_choff + CP), a.Wo, a.Wo, a.Ho, a.B, CP, g.Wp, g.R, &maps.r);
        if (rc != SPK_OK) return rc;
    }
    const long long items = (long long)a.B * g.n_bands * g.P;
    const long long grid = std::min<long long>(items, sm_count());
    const cudaError_t le = launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), (size_t)g.smem_bytes, s, a, g, items, maps);
    if (le != cudaSuccess) {
        set_error("conv_slab4_kernel launch failed: %s", cudaGetErrorString(le));
        return SPK_ERR_CUDA;
    }
    return check_launch("conv_slab4_kernel");
}

}  // namespace

bool conv_slab4_supported(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype) {
    static const bool off = [] { const char *e = getenv("SPK_NO_SLAB4"); return e && e[0] == '1'; }();
    if (off) return false;
    if (in_dtype != SPK_DT_BF16 || out_dtype != SPK_DT_BF16) return false;
    if (a.res != nullptr && res_dtype != SPK_DT_BF16) return false;
    if (a.Cin % 16 || a.Cout % 16 || a.Cin < 16 || a.Cout < 16 || a.Cin > 64 || a.Cout > 64) return false;
    if (a.KH != 3 || a.KW != 3 || a.ph != 1 || a.pw != 1 || a.sh != 1 || a.sw != 1 || a.dh != 1 || a.dw != 1) return false;
    if (a.pro_scale != nullptr || a.gate != nullptr || a.post_scale != nullptr || a.pad_reflect) return false;
    if (a.in_ld % 8 || a.in_choff % 8 || a.out_ld % 8 || a.out_choff % 8) return false;
    if (a.res != nullptr && (a.res_ld % 8 || a.res_choff % 8)) return false;
    if (a.Wo != a.W || a.Ho != a.H || a.W < 8) return false;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.y) & 15) != 0) return false;
    if (a.res != nullptr && (reinterpret_cast<uintptr_t>(a.res) & 15) != 0) return false;
    Slab4Geom g;
    return geometry(a, g);
}

int launch_conv_slab4(const ConvArgs &a, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    Slab4Geom g;
    if (!geometry(a, g)) {
        set_error("conv_slab4: geometry does not fit");
        return SPK_ERR_UNSUPPORTED;
    }
    return chan_pad(a) == 32 ? launch<32>(a, g, s) : launch<64>(a, g, s);
}

}  // namespace spk
