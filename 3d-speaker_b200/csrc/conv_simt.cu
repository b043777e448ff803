// CUDA-core (fp32 FMA) kernels: the exact-fp32 precision mode of the network forward, plus
// the small non-GEMM ops shared by both precision modes (stem conv, CAM gate, statistics
// pooling, AFF blend).  All activations are channels-last.
//
// Reference semantics:
//   conv + BN + act + residual    speakerlab/models/campplus/layers.py:40-67,193-196,209-253
//   CAM context gate              speakerlab/models/campplus/layers.py:93-110
//   statistics pooling            speakerlab/models/campplus/layers.py:26-32,
//                                 speakerlab/models/eres2net/pooling_layers.py:47-55
//   AFF blend                     speakerlab/models/eres2net/fusion.py:22-28
#include <mutex>

#include "ops.cuh"

namespace spk {
namespace {

using bf16 = __nv_bfloat16;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void load(const bf16 *p, float (&v)[4]) {
        const uint2 t = *reinterpret_cast<const uint2 *>(p);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&t.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
    static __device__ __forceinline__ void store(bf16 *p, const float (&v)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 t;
        t.x = *reinterpret_cast<uint32_t *>(&a);
        t.y = *reinterpret_cast<uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(p) = t;
    }
};

// ------------------------------------------------------------------ generic conv (implicit GEMM)
constexpr int BM = 64, BK = 16;

template <typename TIn, typename TOut, typename TRes, int BN>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const ConvArgs a) {
    constexpr int TN = BN / 16;          // columns per thread
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-load assignment: row = tid/4, 4 consecutive k at (tid%4)*4
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const long long lm = m0 + lrow;
    const bool lrow_ok = lm < a.M;
    int lb = 0, lho = 0, lwo = 0;
    if (lrow_ok) {
        const int hw = a.Ho * a.Wo;
        lb = (int)(lm / hw);
        const int r = (int)(lm - (long long)lb * hw);
        lho = r / a.Wo;
        lwo = r - lho * a.Wo;
    }
    const TIn *xin = static_cast<const TIn *>(a.x);
    const float *wgt = static_cast<const float *>(a.w);
    // B-load assignment: BN rows x 16 k, float4 per thread
    const int brow = tid >> 2, bk = (tid & 3) * 4;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.K; k0 += BK) {
        // ---- A tile
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const int tap = k0 / a.Cin;
            const int c = k0 - tap * a.Cin + lk;
            const int kh = tap / a.KW, kw = tap - kh * a.KW;
            const int hi = lho * a.sh - a.ph + kh * a.dh;
            int wi = lwo * a.sw - a.pw + kw * a.dw;
            if (a.pad_reflect) wi = wi < 0 ? -wi : (wi >= a.W ? 2 * (a.W - 1) - wi : wi);
            if (lrow_ok && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) {
                const TIn *src = xin + (((long long)lb * a.H + hi) * a.W + wi) * a.in_ld + a.in_choff + c;
                Vec4<TIn>::load(src, av);
                if (a.pro_scale != nullptr) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        av[i] = fmaf(av[i], __ldg(a.pro_scale + c + i), __ldg(a.pro_shift + c + i));
                        if (a.pro_relu) av[i] = fmaxf(av[i], 0.f);
                    }
                }
            }
        }
        // ---- B tile
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (brow < BN && n0 + brow < a.Cout) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(wgt + (long long)(n0 + brow) * a.K + k0 + bk));
            bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        }
        __syncthreads();   // previous tile fully consumed
#pragma unroll
        for (int i = 0; i < 4; ++i) As[lk + i][lrow] = av[i];
        if (brow < BN) {
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[bk + i][brow] = bv[i];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
            float br[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) br[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }

    // ---- epilogue: affine, residual, activation, gate
    TOut *yout = static_cast<TOut *>(a.y);
    const TRes *res = static_cast<const TRes *>(a.res);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
        const float *grow = nullptr;
        if (a.gate != nullptr) {
            const int hw = a.Ho * a.Wo;
            const int b = (int)(m / hw);
            const int wo = (int)(m % a.Wo);
            grow = a.gate + ((long long)b * a.gate_nwin + wo / a.gate_win) * a.Cout;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= a.Cout) continue;
            float v = acc[i][j];
            if (a.epi_scale != nullptr) v = fmaf(v, __ldg(a.epi_scale + n), __ldg(a.epi_shift + n));
            if (res != nullptr) v += to_f32(res[m * a.res_ld + a.res_choff + n]);
            if (grow != nullptr && a.gate_additive) v += __ldg(grow + n);
            v = apply_act(v, a.act);
            if (grow != nullptr && !a.gate_additive) v *= __ldg(grow + n);
            if (a.post_scale != nullptr) v = apply_act(fmaf(v, __ldg(a.post_scale + n), __ldg(a.post_shift + n)), a.post_act);
            yout[m * a.out_ld + a.out_choff + n] = from_f32<TOut>(v);
        }
    }
}

// ------------------------------------------------------------------ stem: Conv2d(1->Cout,3x3,p=1)+BN+ReLU
// feats [B,T,F] is the image [H=F, W=T] with one channel (DTDNN.py:40-41,112).  One CTA per
// (segment, frequency row): the three input rows are staged in shared memory (transposing the
// [T,F] feature layout on the way), then every thread produces 8 output channels of one pixel so
// consecutive threads write consecutive 16/32-byte pieces of the channels-last output row.
// Pure 32-bit index arithmetic; the output write is the only HBM traffic that matters.
template <typename TOut>
__global__ void __launch_bounds__(256)
stem_kernel(const StemArgs a) {
    extern __shared__ float sm[];
    float *sw = sm;                         // [9][Cout]  (tap-major so a thread reads 8 adjacent channels)
    float *ssc = sw + 9 * a.Cout, *ssh = ssc + a.Cout;
    float *rows = ssh + a.Cout;             // [3][T+2]  zero padded
    const int Tp = a.T + 2;
    const int f = blockIdx.x % a.F, b = blockIdx.x / a.F;
    for (int i = threadIdx.x; i < a.Cout * 9; i += blockDim.x) sw[(i % 9) * a.Cout + i / 9] = a.w[i];
    for (int i = threadIdx.x; i < a.Cout; i += blockDim.x) {
        ssc[i] = a.scale ? a.scale[i] : 1.f;
        ssh[i] = a.shift ? a.shift[i] : 0.f;
    }
    const float *fe = a.feats + (size_t)b * a.T * a.F;
    for (int i = threadIdx.x; i < 3 * Tp; i += blockDim.x) {
        const int kh = i / Tp, tt = i - kh * Tp - 1, ff = f + kh - 1;
        rows[i] = (ff >= 0 && ff < a.F && tt >= 0 && tt < a.T) ? __ldg(fe + (size_t)tt * a.F + ff) : 0.f;
    }
    __syncthreads();
    const int cg = a.Cout / 8;
    TOut *y = static_cast<TOut *>(a.y) + ((size_t)b * a.F + f) * a.T * a.out_ld + a.out_choff;
    for (int idx = threadIdx.x; idx < a.T * cg; idx += blockDim.x) {
        const int t = idx / cg, c0 = (idx - t * cg) * 8;
        float in[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) in[kh * 3 + kw] = rows[kh * Tp + t + kw];
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float4 w0 = *reinterpret_cast<const float4 *>(sw + k * a.Cout + c0);
            const float4 w1 = *reinterpret_cast<const float4 *>(sw + k * a.Cout + c0 + 4);
            o[0] = fmaf(in[k], w0.x, o[0]); o[1] = fmaf(in[k], w0.y, o[1]);
            o[2] = fmaf(in[k], w0.z, o[2]); o[3] = fmaf(in[k], w0.w, o[3]);
            o[4] = fmaf(in[k], w1.x, o[4]); o[5] = fmaf(in[k], w1.y, o[5]);
            o[6] = fmaf(in[k], w1.z, o[6]); o[7] = fmaf(in[k], w1.w, o[7]);
        }
        float r0[4], r1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            r0[j] = apply_act(fmaf(o[j], ssc[c0 + j], ssh[c0 + j]), a.act);
            r1[j] = apply_act(fmaf(o[4 + j], ssc[c0 + 4 + j], ssh[c0 + 4 + j]), a.act);
        }
        TOut *yp = y + (size_t)t * a.out_ld + c0;
        Vec4<TOut>::store(yp, r0);
        Vec4<TOut>::store(yp + 4, r1);
    }
}

// Register-tiled stem for Cout = 8*CG (CG = 4 or 8).  A CTA is one segment x kStemRows frequency rows; the
// kStemRows + 2 input rows are staged once in shared memory (10-18 adjacent floats of every frame: the [T,F] -> [F,T]
// transpose happens here).  A thread owns 8 output channels - their 72 weights and the folded BN live in
// registers - and walks along time: lane = (channel group, pixel), so one warp step writes 32/CG consecutive
// pixels x Cout channels = 512 contiguous bytes (bf16).  Per (pixel, 8 channels): 9 broadcast LDS + 80 FMA.
constexpr int kStemRows = 40;     // frequency rows per CTA (5 per warp): amortises the parameter load, the staging round trip and the barrier
template <typename TOut, int CG>
__global__ void __launch_bounds__(256)
stem_tiled_kernel(const StemArgs a) {
    extern __shared__ float rows[];          // [kStemRows + 2][Tp]  zero padded
    const int Tp = a.T + 2;
    const int fblocks = (a.F + kStemRows - 1) / kStemRows;
    const int fb = blockIdx.x % fblocks, b = blockIdx.x / fblocks;
    const int f0 = fb * kStemRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = lane % CG, p = lane / CG;
    constexpr int PIX = 32 / CG;             // pixels per warp step
    const int c0 = cg * 8;
    // weights as channel pairs: the 72 FMAs of a pixel issue as 36 packed FFMA2 (bit-identical to scalar fmaf)
    // parameters through shared memory: two coalesced loads per thread instead of 88 scalar L2 round trips
    constexpr int COUT = CG * 8;
    __shared__ float s_w[COUT * 9], s_sc[COUT], s_sh[COUT];
    for (int i = threadIdx.x; i < COUT * 9; i += blockDim.x) s_w[i] = __ldg(a.w + i);
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
        s_sc[i] = a.scale ? __ldg(a.scale + i) : 1.f;
        s_sh[i] = a.shift ? __ldg(a.shift + i) : 0.f;
    }
    const float *fe = a.feats + (size_t)b * a.T * a.F;
    constexpr int NR = kStemRows + 2;
    // many loads in flight per thread before the first store (the 1.5 s tile is two rounds of 13): the tile comes
    // from HBM and a CTA cannot start computing until all of it has landed
    constexpr int kLd = 13;
    for (int i0 = threadIdx.x; i0 < NR * Tp; i0 += kLd * blockDim.x) {
        float v[kLd];
#pragma unroll
        for (int u = 0; u < kLd; ++u) {
            const int i = i0 + u * blockDim.x;
            const int tt = i / NR - 1, r = i % NR, ff = f0 + r - 1;     // r fastest: adjacent floats of one frame
            v[u] = (i < NR * Tp && ff >= 0 && ff < a.F && tt >= 0 && tt < a.T) ? __ldg(fe + (size_t)tt * a.F + ff) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kLd; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < NR * Tp) rows[(i % NR) * Tp + i / NR] = v[u];
        }
    }
    __syncthreads();
    float2 w[9][4];
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = s_sc[c0 + j]; sh[j] = s_sh[c0 + j]; }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k][j] = make_float2(s_w[(c0 + 2 * j) * 9 + k], s_w[(c0 + 2 * j + 1) * 9 + k]);
#pragma unroll 1
    for (int rr = 0; rr < kStemRows / 8; ++rr) {
        const int r = warp * (kStemRows / 8) + rr, f = f0 + r;
        if (f >= a.F) break;
        TOut *y = static_cast<TOut *>(a.y) + ((size_t)b * a.F + f) * a.T * a.out_ld + a.out_choff + c0;
        const float *r0 = rows + r * Tp, *r1 = r0 + Tp, *r2 = r1 + Tp;
        for (int t = p; t < a.T; t += PIX) {
            const float in[9] = {r0[t], r0[t + 1], r0[t + 2], r1[t], r1[t + 1], r1[t + 2], r2[t], r2[t + 1], r2[t + 2]};
            float2 o2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o2[j] = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float2 x2 = make_float2(in[k], in[k]);
#pragma unroll
                for (int j = 0; j < 4; ++j) o2[j] = __ffma2_rn(x2, w[k][j], o2[j]);
            }
            float o[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[2 * j] = fmaf(o2[j].x, sc[2 * j], sh[2 * j]);
                o[2 * j + 1] = fmaf(o2[j].y, sc[2 * j + 1], sh[2 * j + 1]);
            }
            apply_act_vec(o, a.act);
            const float q0[4] = {o[0], o[1], o[2], o[3]}, q1[4] = {o[4], o[5], o[6], o[7]};
            TOut *yp = y + (size_t)t * a.out_ld;
            if constexpr (sizeof(TOut) == 2) {      // 8 channels = one 16-byte store; a warp step is 512 contiguous bytes
                __nv_bfloat162 h0 = __floats2bfloat162_rn(q0[0], q0[1]), h1 = __floats2bfloat162_rn(q0[2], q0[3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(q1[0], q1[1]), h3 = __floats2bfloat162_rn(q1[2], q1[3]);
                *reinterpret_cast<uint4 *>(yp) = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1),
                                                            *reinterpret_cast<uint32_t *>(&h2), *reinterpret_cast<uint32_t *>(&h3));
            } else {
                Vec4<TOut>::store(yp, q0);
                Vec4<TOut>::store(yp + 4, q1);
            }
        }
    }
}

// ------------------------------------------------------------------ CAM context gate
// gate[b,w,:] = sigmoid(W2 relu(W1 (mean_T(x) + mean_{window w}(x)) + b1) + b2); one CTA of 256
// threads per segment.  Column sums: thread (g, c4) adds rows g, g+G, ... of 4 adjacent
// channels (8- or 16-byte loads, coalesced across c4), partial sums are combined through shared
// memory in a fixed order (deterministic).  The two tiny mat-vecs use one warp per output.
template <typename TIn>
__global__ void __launch_bounds__(256)
cam_gate_kernel(const CamGateArgs a) {
    extern __shared__ float sh[];
    float *part = sh;                               // [G][C] partial sums of one window
    const int C = a.C, cq = C / 4;
    const int G = blockDim.x / cq;                  // row groups
    float *tot = part + G * C;                      // [C]
    float *win = tot + C;                           // [nwin][C]
    float *ctx = win + a.nwin * C;                  // [C]
    float *hid = ctx + C;                           // [hidden]
    const int b = blockIdx.x;
    const TIn *x = static_cast<const TIn *>(a.x) + (long long)b * a.T * a.in_ld + a.in_choff;
    const int g = threadIdx.x / cq, c4 = (threadIdx.x % cq) * 4;
    for (int c = threadIdx.x; c < C; c += blockDim.x) tot[c] = 0.f;
    for (int w = 0; w < a.nwin; ++w) {
        const int t0 = w * a.seg_len, t1 = min(a.T, t0 + a.seg_len);
        if (g < G) {
            float s[4] = {0.f, 0.f, 0.f, 0.f};
            for (int i = t0 + g; i < t1; i += G) {
                float v[4];
                Vec4<TIn>::load(x + (long long)i * a.in_ld + c4, v);
                s[0] += v[0]; s[1] += v[1]; s[2] += v[2]; s[3] += v[3];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) part[g * C + c4 + q] = s[q];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float s = 0.f;
            for (int gg = 0; gg < G; ++gg) s += part[gg * C + c];
            win[w * C + c] = s / (float)(t1 - t0);
            tot[c] += s;
        }
        __syncthreads();
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) tot[c] = tot[c] / (float)a.T;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int w = 0; w < a.nwin; ++w) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) ctx[c] = a.se_mode ? tot[c] : tot[c] + win[w * C + c];
        __syncthreads();
        for (int j = warp; j < a.hidden; j += nwarps) {
            float s = 0.f;
            for (int c = lane; c < C; c += 32) s = fmaf(__ldg(a.w1 + (long long)j * C + c), ctx[c], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) hid[j] = fmaxf(s + __ldg(a.b1 + j), 0.f);
        }
        __syncthreads();
        for (int o = warp; o < a.Cout; o += nwarps) {
            float s = 0.f;
            for (int j = lane; j < a.hidden; j += 32) s = fmaf(__ldg(a.w2 + (long long)o * a.hidden + j), hid[j], s);
#pragma unroll
            for (int q = 16; q > 0; q >>= 1) s += __shfl_xor_sync(0xffffffffu, s, q);
            if (lane == 0) {
                const float v = s + __ldg(a.b2 + o);
                a.gate[((long long)b * a.nwin + w) * a.Cout + o] = 1.f / (1.f + expf(-v));
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ statistics pooling
// x [B,G,P,ld] -> y[b] = [mean(G,C) | std(G,C)].  Lanes own consecutive channels (coalesced);
// the P positions of one (b,g) are split over the 8 warps of a CTA and combined with the
// pairwise (Chan) mean/M2 update, which is the Welford recurrence for merged partitions.
template <typename TIn>
__global__ void __launch_bounds__(256)
stats_pool_kernel(const StatsPoolArgs a) {
    __shared__ float s_mean[8][32], s_m2[8][32];
    __shared__ int s_n[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int g = blockIdx.y;
    const long long b = blockIdx.z;
    const TIn *x = static_cast<const TIn *>(a.x) + ((b * a.G + g) * a.P) * (long long)a.in_ld + a.in_choff;
    float mean = 0.f, m2 = 0.f;
    int n = 0;
    if (c < a.C) {
        for (int p = warp; p < a.P; p += 8) {
            const float v = to_f32(x[(long long)p * a.in_ld + c]);
            ++n;
            const float d = v - mean;
            mean += d / (float)n;
            m2 = fmaf(d, v - mean, m2);
        }
    } else {
        for (int p = warp; p < a.P; p += 8) ++n;
    }
    s_mean[warp][lane] = mean;
    s_m2[warp][lane] = m2;
    if (lane == 0) s_n[warp] = n;
    __syncthreads();
    if (warp == 0 && c < a.C) {
        float mu = s_mean[0][lane], M2 = s_m2[0][lane];
        int cnt = s_n[0];
        for (int w = 1; w < 8; ++w) {
            const int nb = s_n[w];
            if (nb == 0) continue;
            const float d = s_mean[w][lane] - mu;
            const int tot = cnt + nb;
            mu += d * (float)nb / (float)tot;
            M2 += s_m2[w][lane] + d * d * (float)cnt * (float)nb / (float)tot;
            cnt = tot;
        }
        const float denom = a.unbiased ? (float)(a.P - 1) : (float)a.P;
        const float var = M2 / denom;
        float *y = a.y + b * 2ll * a.G * a.C;
        y[(long long)g * a.C + c] = mu;
        y[(long long)(a.G + g) * a.C + c] = a.var_floor > 0.f ? sqrtf(fmaxf(var, a.var_floor)) : sqrtf(var + a.eps);
    }
}

// Streaming variant: one CTA per (b, g), a thread owns channel PAIRS (8- or 4-byte loads, a position row is one
// contiguous 2*C*sizeof(T)-byte read for the CTA) and walks the P positions once with the shifted-data sums
//   s1 = sum (v - v0),  s2 = sum (v - v0)^2,   v0 = the first position's value
// (no per-element division, and the shift removes the mean^2 >> variance cancellation of plain sum / sum-of-squares):
//   mean = v0 + s1 / P,   M2 = s2 - s1^2 / P.
template <typename TIn>
__global__ void __launch_bounds__(256)
stats_pool_stream_kernel(const StatsPoolArgs a) {
    const int g = blockIdx.x % a.G;
    const long long b = blockIdx.x / a.G;
    const TIn *x = static_cast<const TIn *>(a.x) + ((b * a.G + g) * a.P) * (long long)a.in_ld + a.in_choff;
    float *y = a.y + b * 2ll * a.G * a.C;
    const float inv_p = 1.f / (float)a.P;
    const float denom = a.unbiased ? (float)(a.P - 1) : (float)a.P;
    for (int c = 2 * threadIdx.x; c < a.C; c += 2 * blockDim.x) {
        const TIn *xc = x + c;
        float v0x, v0y;
        if constexpr (sizeof(TIn) == 4) {
            const float2 f = *reinterpret_cast<const float2 *>(xc);
            v0x = f.x; v0y = f.y;
        } else {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(xc);
            v0x = __low2float(h); v0y = __high2float(h);
        }
        float s1x = 0.f, s1y = 0.f, s2x = 0.f, s2y = 0.f;
#pragma unroll 4
        for (int p = 1; p < a.P; ++p) {
            float vx, vy;
            if constexpr (sizeof(TIn) == 4) {
                const float2 f = *reinterpret_cast<const float2 *>(xc + (long long)p * a.in_ld);
                vx = f.x; vy = f.y;
            } else {
                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(xc + (long long)p * a.in_ld);
                vx = __low2float(h); vy = __high2float(h);
            }
            const float dx = vx - v0x, dy = vy - v0y;
            s1x += dx; s1y += dy;
            s2x = fmaf(dx, dx, s2x); s2y = fmaf(dy, dy, s2y);
        }
        const float m2x = fmaxf(s2x - s1x * s1x * inv_p, 0.f), m2y = fmaxf(s2y - s1y * s1y * inv_p, 0.f);
        y[(long long)g * a.C + c] = v0x + s1x * inv_p;
        y[(long long)g * a.C + c + 1] = v0y + s1y * inv_p;
        y[(long long)(a.G + g) * a.C + c] = a.var_floor > 0.f ? sqrtf(fmaxf(m2x / denom, a.var_floor)) : sqrtf(m2x / denom + a.eps);
        y[(long long)(a.G + g) * a.C + c + 1] = a.var_floor > 0.f ? sqrtf(fmaxf(m2y / denom, a.var_floor)) : sqrtf(m2y / denom + a.eps);
    }
}

// ------------------------------------------------------------------ AFF blend
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
aff_blend_kernel(const AffBlendArgs a) {
    const int cg = a.C / 4;
    const long long total = a.M * cg;
    const T *x = static_cast<const T *>(a.x), *y = static_cast<const T *>(a.y), *z = static_cast<const T *>(a.z);
    TO *o = static_cast<TO *>(a.out);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cg) * 4;
        const long long m = idx / cg;
        float xv[4], yv[4], zv[4], ov[4];
        Vec4<T>::load(x + m * a.x_ld + a.x_choff + c, xv);
        if (y == nullptr) {                        // plain (dtype-converting) copy of a channel window
            Vec4<TO>::store(o + m * a.out_ld + a.out_choff + c, xv);
            continue;
        }
        Vec4<T>::load(y + m * a.y_ld + a.y_choff + c, yv);
        if (z != nullptr) {
            Vec4<T>::load(z + m * a.z_ld + a.z_choff + c, zv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float gq = 1.f + tanhf(zv[i]);
                ov[i] = xv[i] * gq + yv[i] * (2.f - gq);
            }
        } else {                                   // plain sum (Res2Net hierarchical add, ERes2NetV2.py:75)
#pragma unroll
            for (int i = 0; i < 4; ++i) ov[i] = xv[i] + yv[i];
        }
        Vec4<TO>::store(o + m * a.out_ld + a.out_choff + c, ov);
    }
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ s, bf16 *__restrict__ d, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        d[i] = __float2bfloat16_rn(s[i]);
}
template <typename T>
__global__ void widen_kernel(const T *__restrict__ s, float *__restrict__ d, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        d[i] = to_f32(s[i]);
}

int grid_for(long long work, int threads) {
    long long g = (work + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

template <typename TIn, typename TOut, typename TRes>
int conv_simt_dispatch(const ConvArgs &a, cudaStream_t s) {
    const long long mt = (a.M + BM - 1) / BM;
    if (mt > 0x7fffffffll) {
        set_error("conv: too many output tiles");
        return SPK_ERR_UNSUPPORTED;
    }
    if (a.Cout <= 32) {
        dim3 grid((unsigned)mt, (a.Cout + 31) / 32);
        conv_simt_kernel<TIn, TOut, TRes, 32><<<grid, 256, 0, s>>>(a);
    } else {
        dim3 grid((unsigned)mt, (a.Cout + 63) / 64);
        conv_simt_kernel<TIn, TOut, TRes, 64><<<grid, 256, 0, s>>>(a);
    }
    return check_launch("conv_simt_kernel");
}

}  // namespace

int launch_conv_simt(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype, cudaStream_t s) {
    if (a.Cin % BK != 0 || a.in_ld % 4 != 0 || a.in_choff % 4 != 0) {
        set_error("conv_simt: Cin=%d must be a multiple of %d and channel offsets 4-aligned", a.Cin, BK);
        return SPK_ERR_UNSUPPORTED;
    }
    if (a.M == 0) return SPK_OK;
    if (linear_supported(a, in_dtype, out_dtype)) return launch_linear(a, in_dtype, out_dtype, s);   // few rows, long K
    const int key = in_dtype * 100 + out_dtype * 10 + (a.res ? res_dtype : out_dtype);
    switch (key) {
        case 0:   return conv_simt_dispatch<float, float, float>(a, s);
        case 11:  return conv_simt_dispatch<float, bf16, bf16>(a, s);
        case 100: return conv_simt_dispatch<bf16, float, float>(a, s);
        case 111: return conv_simt_dispatch<bf16, bf16, bf16>(a, s);
        default:
            set_error("conv_simt: unsupported dtype combination in=%d out=%d res=%d", in_dtype, out_dtype, res_dtype);
            return SPK_ERR_UNSUPPORTED;
    }
}

int launch_stem(const StemArgs &a, int out_dtype, cudaStream_t s) {
    if (a.Cout % 8 != 0 || a.out_ld % 8 != 0 || a.out_choff % 8 != 0) {
        set_error("stem: Cout and channel pitch must be multiples of 8");
        return SPK_ERR_UNSUPPORTED;
    }
    const long long blocks = (long long)a.B * a.F;
    if (blocks == 0) return SPK_OK;
    if (blocks > 0x7fffffffll) {
        set_error("stem: batch too large");
        return SPK_ERR_UNSUPPORTED;
    }
    if ((a.Cout == 32 || a.Cout == 64) && (size_t)(kStemRows + 2) * (a.T + 2) * sizeof(float) <= 200 * 1024) {
        const long long nb = (long long)a.B * ((a.F + kStemRows - 1) / kStemRows);
        const size_t shb = (size_t)(kStemRows + 2) * (a.T + 2) * sizeof(float);
        static std::once_flag once;
        static cudaError_t attr_err = cudaSuccess;
        std::call_once(once, [&] {       // the row tile of a 3 s or 10 s segment needs more than the default 48 KB
            attr_err = cudaFuncSetAttribute(stem_tiled_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(stem_tiled_kernel<bf16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(stem_tiled_kernel<float, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(stem_tiled_kernel<bf16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        });
        if (attr_err != cudaSuccess) {
            set_error("cudaFuncSetAttribute(stem_tiled) failed: %s", cudaGetErrorString(attr_err));
            return SPK_ERR_CUDA;
        }
        if (a.Cout == 32) {
            if (out_dtype == SPK_DT_F32) stem_tiled_kernel<float, 4><<<(unsigned)nb, 256, shb, s>>>(a);
            else stem_tiled_kernel<bf16, 4><<<(unsigned)nb, 256, shb, s>>>(a);
        } else {
            if (out_dtype == SPK_DT_F32) stem_tiled_kernel<float, 8><<<(unsigned)nb, 256, shb, s>>>(a);
            else stem_tiled_kernel<bf16, 8><<<(unsigned)nb, 256, shb, s>>>(a);
        }
        return check_launch("stem_tiled_kernel");
    }
    const size_t sh = ((size_t)a.Cout * 11 + 3 * (a.T + 2)) * sizeof(float);
    if (sh > 48 * 1024) {
        set_error("stem: %d frames exceed the shared-memory row buffer", a.T);
        return SPK_ERR_UNSUPPORTED;
    }
    if (out_dtype == SPK_DT_F32) stem_kernel<float><<<(unsigned)blocks, 256, sh, s>>>(a);
    else stem_kernel<bf16><<<(unsigned)blocks, 256, sh, s>>>(a);
    return check_launch("stem_kernel");
}

int launch_cam_gate(const CamGateArgs &a, int in_dtype, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    if (se_gate_cluster_supported(a, in_dtype)) return launch_se_gate_cluster(a, in_dtype, s);
    const int threads = 256;
    if (a.C % 4 != 0 || a.C / 4 > threads || a.in_ld % 4 != 0 || a.in_choff % 4 != 0) {
        set_error("cam_gate: C=%d must be a multiple of 4 and <= %d", a.C, threads * 4);
        return SPK_ERR_UNSUPPORTED;
    }
    const int G = threads / (a.C / 4);
    const size_t sh = ((size_t)a.C * (G + a.nwin + 2) + a.hidden) * sizeof(float);
    if (sh > 48 * 1024) {
        set_error("cam_gate: %d windows x %d channels exceed shared memory", a.nwin, a.C);
        return SPK_ERR_UNSUPPORTED;
    }
    if (in_dtype == SPK_DT_F32) cam_gate_kernel<float><<<a.B, threads, sh, s>>>(a);
    else cam_gate_kernel<bf16><<<a.B, threads, sh, s>>>(a);
    return check_launch("cam_gate_kernel");
}

int launch_stats_pool(const StatsPoolArgs &a, int in_dtype, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    if (a.G > 65535 || a.B > 65535) {
        set_error("stats_pool: grid too large (G=%d, B=%d)", a.G, a.B);
        return SPK_ERR_UNSUPPORTED;
    }
    if (stats_pool_sliced_supported(a)) return launch_stats_pool_sliced(a, in_dtype, s);     // long axes (10 s chunks)
    if (a.C % 2 == 0 && a.in_ld % 2 == 0 && a.in_choff % 2 == 0 && (long long)a.B * a.G < 0x7fffffffll &&
        (reinterpret_cast<uintptr_t>(a.x) & 7) == 0) {
        const unsigned blocks = (unsigned)((long long)a.B * a.G);
        if (in_dtype == SPK_DT_F32) stats_pool_stream_kernel<float><<<blocks, 256, 0, s>>>(a);
        else stats_pool_stream_kernel<bf16><<<blocks, 256, 0, s>>>(a);
        return check_launch("stats_pool_stream_kernel");
    }
    dim3 grid((a.C + 31) / 32, a.G, a.B);
    if (in_dtype == SPK_DT_F32) stats_pool_kernel<float><<<grid, 256, 0, s>>>(a);
    else stats_pool_kernel<bf16><<<grid, 256, 0, s>>>(a);
    return check_launch("stats_pool_kernel");
}

int launch_aff_blend(const AffBlendArgs &a, int dtype, int out_dtype, cudaStream_t s) {
    if (a.C % 4 != 0) {
        set_error("aff_blend: C must be a multiple of 4");
        return SPK_ERR_UNSUPPORTED;
    }
    const long long work = a.M * (a.C / 4);
    if (work == 0) return SPK_OK;
    const int g = grid_for(work, 256);
    if (dtype == SPK_DT_F32 && out_dtype == SPK_DT_F32) aff_blend_kernel<float, float><<<g, 256, 0, s>>>(a);
    else if (dtype == SPK_DT_BF16 && out_dtype == SPK_DT_BF16) aff_blend_kernel<bf16, bf16><<<g, 256, 0, s>>>(a);
    else if (dtype == SPK_DT_BF16 && out_dtype == SPK_DT_F32) aff_blend_kernel<bf16, float><<<g, 256, 0, s>>>(a);
    else if (dtype == SPK_DT_F32 && out_dtype == SPK_DT_BF16 && a.y == nullptr) aff_blend_kernel<float, bf16><<<g, 256, 0, s>>>(a);
    else {
        set_error("aff_blend: unsupported dtype combination");
        return SPK_ERR_UNSUPPORTED;
    }
    return check_launch("aff_blend_kernel");
}

int launch_f32_to_bf16(const float *src, __nv_bfloat16 *dst, long long n, cudaStream_t s) {
    if (n == 0) return SPK_OK;
    f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, s>>>(src, dst, n);
    return check_launch("f32_to_bf16_kernel");
}

int launch_widen(const void *src, int dtype, float *dst, long long n, cudaStream_t s) {
    if (n == 0) return SPK_OK;
    if (dtype == SPK_DT_F32) widen_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const float *>(src), dst, n);
    else widen_kernel<bf16><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const bf16 *>(src), dst, n);
    return check_launch("widen_kernel");
}

}  // namespace spk
