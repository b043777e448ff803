// Latency-shaped ops: few rows or few segments, long reductions.  The generic kernels give every output tile (or
// every segment) to ONE CTA, which leaves most of the chip idle exactly where the reduction is longest:
//   * the embedding layers (seg_1 / dense / fc: [B, K] x [N, K]^T with K = 1,024 ... 20,480 and B a few hundred rows;
//     speakerlab/models/eres2net/ERes2NetV2.py:251, campplus/DTDNN.py:102-103, ecapa_tdnn/ECAPA_TDNN.py:470) ran as 6
//     CTAs of the implicit-GEMM kernel: 1.6 ms per call, 7-13 % of an ERes2NetV2 / ECAPA forward;
//   * ECAPA's squeeze-excitation context (mean over ~1,000 frames of 1,024 channels, ECAPA_TDNN.py:203-222), the
//     statistics of its attentive pooling (:257-285) and the global-context statistics in front of it.
// Here the long axis is split over a thread-block CLUSTER (split-K / split-T) and the partial results are combined
// through distributed shared memory in rank order, so results do not depend on the launch shape and need no scratch
// buffer or atomics; the pooling kernels split positions over the warps of a CTA and channels over the grid.
#include <cooperative_groups.h>

#include "ops.cuh"

namespace cg = cooperative_groups;

namespace spk {
namespace {

using bf16 = __nv_bfloat16;

template <typename T> __device__ __forceinline__ void ld4(const T *p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float *p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<bf16>(const bf16 *p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2 *>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&t.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <typename T> __device__ __forceinline__ float2 ld2(const T *p);
template <> __device__ __forceinline__ float2 ld2<float>(const float *p) { return *reinterpret_cast<const float2 *>(p); }
template <> __device__ __forceinline__ float2 ld2<bf16>(const bf16 *p) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(p);
    return make_float2(__low2float(h), __high2float(h));
}

// ------------------------------------------------------------------ split-K linear layer
// y[m, n] = epilogue(sum_k f(x[m, k]) w[n, k]); CTA tile 32 x 32, thread 2 x 2, K in steps of 32 through shared memory.
// grid (M tiles, N tiles, S), cluster (1, 1, S): rank z reduces K range z; rank 0 adds the S partial tiles in rank
// order and runs the epilogue.  S depends on K only, so a row's result does not depend on the batch it arrived in.
constexpr int LBM = 32, LBN = 32, LBK = 32;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
linear_splitk_kernel(const ConvArgs a, int kper) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ __align__(16) float As[LBK][LBM + 4];
    __shared__ __align__(16) float Bs[LBK][LBN + 4];
    __shared__ __align__(16) float red[LBM][LBN + 2];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * LBM;
    const int n0 = blockIdx.y * LBN;
    const int rank = (int)cluster.block_rank(), nrank = (int)cluster.num_blocks();
    const int kbeg = rank * kper, kend = min(a.K, kbeg + kper);
    const int lrow = tid >> 3, k4 = (tid & 7) * 4;
    const long long lm = m0 + lrow;
    const TIn *xrow = static_cast<const TIn *>(a.x) + lm * a.in_ld + a.in_choff;
    const float *wrow = static_cast<const float *>(a.w) + (long long)(n0 + lrow) * a.K;
    const bool a_ok = lm < a.M, b_ok = n0 + lrow < a.Cout;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    // the next step's global loads are issued before the current step's 32 x 4 FMAs (register double buffering)
    auto fetch = [&](int k0, float (&av)[4], float (&bv)[4]) {
        const int k = k0 + k4;
#pragma unroll
        for (int i = 0; i < 4; ++i) { av[i] = 0.f; bv[i] = 0.f; }
        if (a_ok && k < kend) {
            ld4<TIn>(xrow + k, av);
            if (a.pro_scale != nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    av[i] = fmaf(av[i], __ldg(a.pro_scale + k + i), __ldg(a.pro_shift + k + i));
                    if (a.pro_relu) av[i] = fmaxf(av[i], 0.f);
                }
            }
        }
        if (b_ok && k < kend) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(wrow + k));
            bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        }
    };
    float av[4], bv[4], an[4], bn[4];
    if (kbeg < kend) fetch(kbeg, av, bv);
    for (int k0 = kbeg; k0 < kend; k0 += LBK) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[k4 + i][lrow] = av[i];
            Bs[k4 + i][lrow] = bv[i];
        }
        __syncthreads();
        if (k0 + LBK < kend) fetch(k0 + LBK, an, bn);
#pragma unroll
        for (int kk = 0; kk < LBK; ++kk) {
            const float2 a2 = *reinterpret_cast<const float2 *>(&As[kk][ty * 2]);
            const float2 b2 = *reinterpret_cast<const float2 *>(&Bs[kk][tx * 2]);
            acc[0][0] = fmaf(a2.x, b2.x, acc[0][0]);
            acc[0][1] = fmaf(a2.x, b2.y, acc[0][1]);
            acc[1][0] = fmaf(a2.y, b2.x, acc[1][0]);
            acc[1][1] = fmaf(a2.y, b2.y, acc[1][1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { av[i] = an[i]; bv[i] = bn[i]; }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) red[ty * 2 + i][tx * 2 + j] = acc[i][j];
    cluster.sync();
    if (rank == 0) {
        for (int r = 1; r < nrank; ++r) {
            const float *peer = cluster.map_shared_rank(&red[0][0], r);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j] += peer[(ty * 2 + i) * (LBN + 2) + tx * 2 + j];
        }
        TOut *yout = static_cast<TOut *>(a.y);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const long long m = m0 + ty * 2 + i;
            if (m >= a.M) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = n0 + tx * 2 + j;
                if (n >= a.Cout) continue;
                float v = acc[i][j];
                if (a.epi_scale != nullptr) v = fmaf(v, __ldg(a.epi_scale + n), __ldg(a.epi_shift + n));
                v = apply_act(v, a.act);
                if (a.post_scale != nullptr) v = apply_act(fmaf(v, __ldg(a.post_scale + n), __ldg(a.post_shift + n)), a.post_act);
                yout[m * a.out_ld + a.out_choff + n] = from_f32<TOut>(v);
            }
        }
    }
    cluster.sync();      // peers keep their shared memory alive until rank 0 has read it
}

// ------------------------------------------------------------------ squeeze-excitation gate, split over time
// gate[b, :] = sigmoid(W2 relu(W1 mean_t x[b, t, :] + b1) + b2).  A cluster of 8 CTAs per segment: rank r sums frames
// [r T/8, (r+1) T/8) (16-byte loads, row groups combined in a fixed order), every rank then adds the 8 partial sums in
// rank order, computes hidden/8 of the first layer, reads the other ranks' hidden units through distributed shared
// memory and writes Cout/8 of the gate.
constexpr int kSeRanks = 8;

template <typename TIn>
__global__ void __launch_bounds__(256)
se_gate_cluster_kernel(const CamGateArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int CV = 16 / (int)sizeof(TIn);          // channels per 16-byte load
    extern __shared__ __align__(16) float sh[];
    const int C = a.C, cq = C / CV, G = blockDim.x / cq;
    float *part = sh;                 // [G][C]
    float *csum = part + G * C;       // [C]       this rank's column sums (peers read it)
    float *ctx = csum + C;            // [C]
    float *hloc = ctx + C;            // [hidden]  this rank's slice of the hidden layer (peers read it)
    float *hid = hloc + a.hidden;     // [hidden]
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.y;
    const int per = (a.T + kSeRanks - 1) / kSeRanks;
    const int t0 = rank * per, t1 = min(a.T, t0 + per);
    const TIn *x = static_cast<const TIn *>(a.x) + (long long)b * a.T * a.in_ld + a.in_choff;
    const int g = threadIdx.x / cq, c0 = (threadIdx.x % cq) * CV;
    if (g < G) {
        float s[CV];
#pragma unroll
        for (int q = 0; q < CV; ++q) s[q] = 0.f;
#pragma unroll 4
        for (int t = t0 + g; t < t1; t += G) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(x + (long long)t * a.in_ld + c0);
            if constexpr (sizeof(TIn) == 4) {
                s[0] += __uint_as_float(raw.x); s[1] += __uint_as_float(raw.y);
                s[2] += __uint_as_float(raw.z); s[3] += __uint_as_float(raw.w);
            } else {
                const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w4[q]);
                    s[2 * q] += __low2float(h);
                    s[2 * q + 1] += __high2float(h);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < CV; ++q) part[g * C + c0 + q] = s[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int gg = 0; gg < G; ++gg) s += part[gg * C + c];
        csum[c] = s;
    }
    cluster.sync();
    const float inv_t = 1.f / (float)a.T;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < kSeRanks; ++r) s += cluster.map_shared_rank(csum, r)[c];
        ctx[c] = s * inv_t;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int hper = (a.hidden + kSeRanks - 1) / kSeRanks;
    for (int j = rank * hper + warp; j < min(a.hidden, (rank + 1) * hper); j += nwarps) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(a.w1 + (long long)j * C + c), ctx[c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) hloc[j] = fmaxf(s + __ldg(a.b1 + j), 0.f);
    }
    cluster.sync();
    for (int j = threadIdx.x; j < a.hidden; j += blockDim.x) hid[j] = cluster.map_shared_rank(hloc, j / hper)[j];
    __syncthreads();
    const int oper = (a.Cout + kSeRanks - 1) / kSeRanks;
    for (int o = rank * oper + warp; o < min(a.Cout, (rank + 1) * oper); o += nwarps) {
        float s = 0.f;
        for (int j = lane; j < a.hidden; j += 32) s = fmaf(__ldg(a.w2 + (long long)o * a.hidden + j), hid[j], s);
#pragma unroll
        for (int q = 16; q > 0; q >>= 1) s += __shfl_xor_sync(0xffffffffu, s, q);
        if (lane == 0) a.gate[(long long)b * a.Cout + o] = 1.f / (1.f + expf(-(s + __ldg(a.b2 + o))));
    }
    cluster.sync();      // nobody leaves while a peer may still read hloc
}

// ------------------------------------------------------------------ statistics pooling over long axes
// CTA = 64 channels x 8 position slices (warp w walks p = 1 + w, 9 + w, ...; lanes own channel pairs).  Same shifted
// sums as stats_pool_stream_kernel (shift = the value at position 0), slices added in warp order.
template <typename TIn>
__global__ void __launch_bounds__(256)
stats_pool_sliced_kernel(const StatsPoolArgs a) {
    __shared__ float2 s1s[8][32], s2s[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x % a.G;
    const long long b = blockIdx.x / a.G;
    const int c = blockIdx.y * 64 + 2 * lane;
    const bool ok = c < a.C;
    const TIn *x = static_cast<const TIn *>(a.x) + ((b * a.G + g) * a.P) * (long long)a.in_ld + a.in_choff + (ok ? c : 0);
    const float2 v0 = ld2<TIn>(x);
    float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
    if (ok) {
#pragma unroll 4
        for (int p = 1 + warp; p < a.P; p += 8) {
            const float2 v = ld2<TIn>(x + (long long)p * a.in_ld);
            const float dx = v.x - v0.x, dy = v.y - v0.y;
            s1.x += dx; s1.y += dy;
            s2.x = fmaf(dx, dx, s2.x); s2.y = fmaf(dy, dy, s2.y);
        }
    }
    s1s[warp][lane] = s1;
    s2s[warp][lane] = s2;
    __syncthreads();
    if (warp == 0 && ok) {
        for (int w = 1; w < 8; ++w) {
            s1.x += s1s[w][lane].x; s1.y += s1s[w][lane].y;
            s2.x += s2s[w][lane].x; s2.y += s2s[w][lane].y;
        }
        const float inv_p = 1.f / (float)a.P;
        const float denom = a.unbiased ? (float)(a.P - 1) : (float)a.P;
        const float m2x = fmaxf(s2.x - s1.x * s1.x * inv_p, 0.f), m2y = fmaxf(s2.y - s1.y * s1.y * inv_p, 0.f);
        float *y = a.y + b * 2ll * a.G * a.C;
        y[(long long)g * a.C + c] = v0.x + s1.x * inv_p;
        y[(long long)g * a.C + c + 1] = v0.y + s1.y * inv_p;
        y[(long long)(a.G + g) * a.C + c] = a.var_floor > 0.f ? sqrtf(fmaxf(m2x / denom, a.var_floor)) : sqrtf(m2x / denom + a.eps);
        y[(long long)(a.G + g) * a.C + c + 1] = a.var_floor > 0.f ? sqrtf(fmaxf(m2y / denom, a.var_floor)) : sqrtf(m2y / denom + a.eps);
    }
}

// ------------------------------------------------------------------ attentive statistics, one pass
// Online softmax over positions with the weighted first and second moments of (x - x[0]) carried along: state
// (m, den, S1, S2) per channel; a larger maximum rescales the three sums by exp(m_old - m_new).  Same 64-channel x
// 8-slice CTA shape; the 8 slice states are merged in warp order.  One read of the logits and of x (the three-pass
// kernel read 5x as much).
struct AspState { float m, den, s1, s2; };
__device__ __forceinline__ void asp_step(AspState &st, float l, float d) {
    if (l > st.m) {
        const float f = expf(st.m - l);       // exp(-inf) = 0 on the first step
        st.den *= f; st.s1 *= f; st.s2 *= f;
        st.m = l;
    }
    const float e = expf(l - st.m);
    st.den += e;
    st.s1 = fmaf(e, d, st.s1);
    st.s2 = fmaf(e * d, d, st.s2);
}
__device__ __forceinline__ void asp_merge(AspState &st, const AspState &o) {
    if (o.den == 0.f) return;
    const float m = fmaxf(st.m, o.m);
    const float fa = st.den == 0.f ? 0.f : expf(st.m - m), fb = expf(o.m - m);
    st.den = st.den * fa + o.den * fb;
    st.s1 = st.s1 * fa + o.s1 * fb;
    st.s2 = st.s2 * fa + o.s2 * fb;
    st.m = m;
}

template <typename TL, typename TX>
__global__ void __launch_bounds__(256)
asp_pool_online_kernel(const AspPoolArgs a) {
    __shared__ AspState sst[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long b = blockIdx.y;
    const int c = blockIdx.x * 64 + 2 * lane;
    const bool ok = c < a.C;
    const TL *l = static_cast<const TL *>(a.logits) + b * a.P * (long long)a.l_ld + a.l_choff + (ok ? c : 0);
    const TX *x = static_cast<const TX *>(a.x) + b * a.P * (long long)a.x_ld + a.x_choff + (ok ? c : 0);
    const float2 x0 = ld2<TX>(x);
    AspState sa{-INFINITY, 0.f, 0.f, 0.f}, sb{-INFINITY, 0.f, 0.f, 0.f};
    if (ok) {
#pragma unroll 4
        for (int p = warp; p < a.P; p += 8) {
            const float2 lv = ld2<TL>(l + (long long)p * a.l_ld);
            const float2 xv = ld2<TX>(x + (long long)p * a.x_ld);
            asp_step(sa, lv.x, xv.x - x0.x);
            asp_step(sb, lv.y, xv.y - x0.y);
        }
    }
    sst[warp][2 * lane] = sa;
    sst[warp][2 * lane + 1] = sb;
    __syncthreads();
    if (warp == 0 && ok) {
        for (int w = 1; w < 8; ++w) {
            asp_merge(sa, sst[w][2 * lane]);
            asp_merge(sb, sst[w][2 * lane + 1]);
        }
        float *y = a.out + b * 2ll * a.C;
        const float ia = 1.f / sa.den, ib = 1.f / sb.den;
        const float ma = sa.s1 * ia, mb = sb.s1 * ib;
        y[c] = x0.x + ma;
        y[c + 1] = x0.y + mb;
        y[a.C + c] = sqrtf(fmaxf(sa.s2 * ia - ma * ma, a.var_floor));
        y[a.C + c + 1] = sqrtf(fmaxf(sb.s2 * ib - mb * mb, a.var_floor));
    }
}

// ------------------------------------------------------------------ reflect-padding edge fix
// A 1-D 'same' conv with F.pad(mode='reflect') (ECAPA-TDNN's Conv1d, ECAPA_TDNN.py:49-77) differs from the zero-padded
// conv only at the pw positions next to each end of a segment.  The tensor-core path runs the zero-padded conv with TMA
// im2col loads; this kernel then recomputes those 2 pw positions per segment with mirrored taps and the full epilogue.
// grid (B, 2 ends), thread = output channel: the (KW - 1) dw + 1 input rows an end needs are staged in shared memory
// (every thread reads the same element: broadcast), each thread walks its own weight row once for all edge positions.
constexpr int kFixMaxEdge = 8;

__device__ __forceinline__ void load8_f32(const bf16 *p, float (&v)[8]) {
    const uint4 raw = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w4[q]);
        v[2 * q] = __low2float(h);
        v[2 * q + 1] = __high2float(h);
    }
}
__device__ __forceinline__ void load8_f32(const float *p, float (&v)[8]) {
    const float4 lo = *reinterpret_cast<const float4 *>(p), hi = *reinterpret_cast<const float4 *>(p + 4);
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

// TIn: type of the activations AND of the weights (bf16 mode / fp32 mode)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
reflect_edge_fix_kernel(const ConvArgs a) {
    extern __shared__ float xs[];                       // [rows][Cin]
    const int b = blockIdx.x, right = blockIdx.y;
    const int E = a.pw, span = (a.KW - 1) * a.dw + 1;    // edge positions per end, input rows they touch
    const int row0 = right ? a.W - span : 0;
    const TIn *x = static_cast<const TIn *>(a.x) + ((long long)b * a.W + row0) * a.in_ld + a.in_choff;
    for (int idx = threadIdx.x; idx < span * a.Cin; idx += blockDim.x) {
        const int r = idx / a.Cin, c = idx - r * a.Cin;
        xs[idx] = to_f32(x[(long long)r * a.in_ld + c]);
    }
    __syncthreads();
    const TIn *w = static_cast<const TIn *>(a.w);        // [Cout][KW][Cin]
    TOut *y = static_cast<TOut *>(a.y);
    for (int n = threadIdx.x; n < a.Cout; n += blockDim.x) {
        float acc[kFixMaxEdge];
#pragma unroll
        for (int e = 0; e < kFixMaxEdge; ++e) acc[e] = 0.f;
        for (int j = 0; j < a.KW; ++j) {
            int rows[kFixMaxEdge];                        // staged row read by edge position e through tap j
#pragma unroll
            for (int e = 0; e < kFixMaxEdge; ++e) {
                const int t = right ? a.W - E + e : e;
                int wi = t - a.pw + j * a.dw;
                wi = wi < 0 ? -wi : (wi >= a.W ? 2 * (a.W - 1) - wi : wi);
                rows[e] = (wi - row0) * a.Cin;
            }
            const TIn *wr = w + ((long long)n * a.KW + j) * a.Cin;
            for (int c = 0; c < a.Cin; c += 8) {
                float wv[8];
                load8_f32(wr + c, wv);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
#pragma unroll
                    for (int e = 0; e < kFixMaxEdge; ++e)
                        if (e < E) acc[e] = fmaf(wv[q], xs[rows[e] + c + q], acc[e]);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < kFixMaxEdge; ++e) {
            if (e >= E) break;
            const int t = right ? a.W - E + e : e;
            float v = acc[e];
            if (a.epi_scale != nullptr) v = fmaf(v, __ldg(a.epi_scale + n), __ldg(a.epi_shift + n));
            v = apply_act(v, a.act);
            if (a.post_scale != nullptr) v = apply_act(fmaf(v, __ldg(a.post_scale + n), __ldg(a.post_shift + n)), a.post_act);
            y[((long long)b * a.W + t) * a.out_ld + a.out_choff + n] = from_f32<TOut>(v);
        }
    }
}

template <typename K, typename... Args>
cudaError_t launch_cluster(K kern, dim3 grid, dim3 block, size_t smem, dim3 cluster, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster.x; attr[0].val.clusterDim.y = cluster.y; attr[0].val.clusterDim.z = cluster.z;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

int linear_split(int K) {
    int s = K / 1024;
    return s < 1 ? 1 : (s > 8 ? 8 : s);
}

template <typename TIn, typename TOut>
int linear_dispatch(const ConvArgs &a, cudaStream_t s) {
    const int S = linear_split(a.K);
    const int kper = ((a.K + S - 1) / S + LBK - 1) / LBK * LBK;
    dim3 grid((unsigned)((a.M + LBM - 1) / LBM), (unsigned)((a.Cout + LBN - 1) / LBN), (unsigned)S);
    const cudaError_t e = launch_cluster(linear_splitk_kernel<TIn, TOut>, grid, dim3(256), 0, dim3(1, 1, S), s, a, kper);
    if (e != cudaSuccess) {
        set_error("linear_splitk_kernel launch failed: %s", cudaGetErrorString(e));
        return SPK_ERR_CUDA;
    }
    return check_launch("linear_splitk_kernel");
}

}  // namespace

bool linear_supported(const ConvArgs &a, int in_dtype, int out_dtype) {
    if (a.KH != 1 || a.KW != 1 || a.sh != 1 || a.sw != 1 || a.ph != 0 || a.pw != 0) return false;
    if (a.gate != nullptr || a.res != nullptr) return false;
    if (a.M > 4096 || a.K < 512 || a.K % 4 != 0 || a.in_ld % 4 != 0 || a.in_choff % 4 != 0) return false;
    if (out_dtype != SPK_DT_F32 && out_dtype != SPK_DT_BF16) return false;
    return in_dtype == SPK_DT_F32 || in_dtype == SPK_DT_BF16;
}

int launch_linear(const ConvArgs &a, int in_dtype, int out_dtype, cudaStream_t s) {
    if (a.M == 0) return SPK_OK;
    if (in_dtype == SPK_DT_F32 && out_dtype == SPK_DT_F32) return linear_dispatch<float, float>(a, s);
    if (in_dtype == SPK_DT_F32 && out_dtype == SPK_DT_BF16) return linear_dispatch<float, bf16>(a, s);
    if (in_dtype == SPK_DT_BF16 && out_dtype == SPK_DT_F32) return linear_dispatch<bf16, float>(a, s);
    return linear_dispatch<bf16, bf16>(a, s);
}

bool reflect_edge_fix_supported(const ConvArgs &a) {
    if (!a.pad_reflect || a.H != 1 || a.KH != 1 || a.sw != 1 || a.Wo != a.W || a.B > 65535) return false;
    if (a.pw != a.dw * (a.KW - 1) / 2 || a.pw < 1 || a.pw > kFixMaxEdge || (a.KW - 1) * a.dw + 1 > a.W || 2 * a.pw > a.W) return false;
    if (a.gate != nullptr || a.res != nullptr || a.pro_scale != nullptr || a.Cin % 8 != 0) return false;
    return (size_t)((a.KW - 1) * a.dw + 1) * a.Cin * sizeof(float) <= 48 * 1024;
}

int launch_reflect_edge_fix(const ConvArgs &a, int in_dtype, int out_dtype, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    const size_t sh = (size_t)((a.KW - 1) * a.dw + 1) * a.Cin * sizeof(float);
    const dim3 grid((unsigned)a.B, 2);
    if (in_dtype == SPK_DT_F32) {
        if (out_dtype != SPK_DT_F32) {
            set_error("reflect_edge_fix: fp32 activations take an fp32 output");
            return SPK_ERR_UNSUPPORTED;
        }
        reflect_edge_fix_kernel<float, float><<<grid, 256, sh, s>>>(a);
    } else if (out_dtype == SPK_DT_BF16) reflect_edge_fix_kernel<bf16, bf16><<<grid, 256, sh, s>>>(a);
    else reflect_edge_fix_kernel<bf16, float><<<grid, 256, sh, s>>>(a);
    return check_launch("reflect_edge_fix_kernel");
}

bool se_gate_cluster_supported(const CamGateArgs &a, int in_dtype) {
    const int cv = in_dtype == SPK_DT_F32 ? 4 : 8;
    if (!a.se_mode || a.nwin != 1) return false;
    if (a.C % cv != 0 || a.C / cv > 256 || a.in_ld % cv != 0 || a.in_choff % cv != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0) return false;
    return a.T >= 64 && a.B <= 65535;
}

int launch_se_gate_cluster(const CamGateArgs &a, int in_dtype, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    const int cv = in_dtype == SPK_DT_F32 ? 4 : 8;
    const int G = 256 / (a.C / cv);
    const size_t sh = ((size_t)a.C * (G + 2) + 2 * (size_t)a.hidden) * sizeof(float);
    if (sh > 48 * 1024) {
        set_error("se_gate: %d channels exceed shared memory", a.C);
        return SPK_ERR_UNSUPPORTED;
    }
    const dim3 grid(kSeRanks, (unsigned)a.B), cl(kSeRanks, 1, 1);
    const cudaError_t e = in_dtype == SPK_DT_F32 ? launch_cluster(se_gate_cluster_kernel<float>, grid, dim3(256), sh, cl, s, a)
                                                 : launch_cluster(se_gate_cluster_kernel<bf16>, grid, dim3(256), sh, cl, s, a);
    if (e != cudaSuccess) {
        set_error("se_gate_cluster_kernel launch failed: %s", cudaGetErrorString(e));
        return SPK_ERR_CUDA;
    }
    return check_launch("se_gate_cluster_kernel");
}

bool stats_pool_sliced_supported(const StatsPoolArgs &a) {
    return a.P >= 256 && a.C % 2 == 0 && a.in_ld % 2 == 0 && a.in_choff % 2 == 0 && (long long)a.B * a.G < 0x7fffffffll &&
           (reinterpret_cast<uintptr_t>(a.x) & 7) == 0;
}

int launch_stats_pool_sliced(const StatsPoolArgs &a, int in_dtype, cudaStream_t s) {
    const dim3 grid((unsigned)((long long)a.B * a.G), (unsigned)((a.C + 63) / 64));
    if (in_dtype == SPK_DT_F32) stats_pool_sliced_kernel<float><<<grid, 256, 0, s>>>(a);
    else stats_pool_sliced_kernel<bf16><<<grid, 256, 0, s>>>(a);
    return check_launch("stats_pool_sliced_kernel");
}

bool asp_pool_online_supported(const AspPoolArgs &a) {
    return a.C % 2 == 0 && a.l_ld % 2 == 0 && a.l_choff % 2 == 0 && a.x_ld % 2 == 0 && a.x_choff % 2 == 0 && a.B <= 65535 &&
           (reinterpret_cast<uintptr_t>(a.logits) & 7) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 7) == 0;
}

int launch_asp_pool_online(const AspPoolArgs &a, int l_dtype, int x_dtype, cudaStream_t s) {
    const dim3 grid((unsigned)((a.C + 63) / 64), (unsigned)a.B);
    if (l_dtype == SPK_DT_F32 && x_dtype == SPK_DT_F32) asp_pool_online_kernel<float, float><<<grid, 256, 0, s>>>(a);
    else if (l_dtype == SPK_DT_BF16 && x_dtype == SPK_DT_BF16) asp_pool_online_kernel<bf16, bf16><<<grid, 256, 0, s>>>(a);
    else if (l_dtype == SPK_DT_F32 && x_dtype == SPK_DT_BF16) asp_pool_online_kernel<float, bf16><<<grid, 256, 0, s>>>(a);
    else {
        set_error("asp_pool: unsupported dtype combination %d/%d", l_dtype, x_dtype);
        return SPK_ERR_UNSUPPORTED;
    }
    return check_launch("asp_pool_online_kernel");
}

}  // namespace spk
