// Small memory-bound ops of ECAPA-TDNN (speakerlab/models/ecapa_tdnn/ECAPA_TDNN.py) that are not convolutions:
// squeeze-excitation scaling with the block residual (:222, :345) and the softmax-weighted statistics of the
// attentive pooling layer (:279-285).  Channels-last, lanes own consecutive channels.
#include "ops.cuh"

namespace spk {
namespace {

using bf16 = __nv_bfloat16;

template <typename T> __device__ __forceinline__ void load4(const T *p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float *p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16 *p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2 *>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&t.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <typename T> __device__ __forceinline__ void store4(T *p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<bf16>(bf16 *p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t *>(&a);
    t.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = t;
}

template <typename T, typename TR, typename TO>
__global__ void __launch_bounds__(256)
se_scale_kernel(const SeScaleArgs a) {
    const int cg = a.C / 4;
    const long long total = a.B * a.P * cg;
    const T *x = static_cast<const T *>(a.x);
    const TR *res = static_cast<const TR *>(a.res);
    TO *o = static_cast<TO *>(a.out);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cg) * 4;
        const long long m = idx / cg;
        const long long b = m / a.P;
        float xv[4], rv[4] = {0.f, 0.f, 0.f, 0.f};
        load4<T>(x + m * a.x_ld + a.x_choff + c, xv);
        if (res != nullptr) load4<TR>(res + m * a.res_ld + a.res_choff + c, rv);
        const float4 g = __ldg(reinterpret_cast<const float4 *>(a.gate + b * a.C + c));
        const float ov[4] = {fmaf(xv[0], g.x, rv[0]), fmaf(xv[1], g.y, rv[1]), fmaf(xv[2], g.z, rv[2]), fmaf(xv[3], g.w, rv[3])};
        store4<TO>(o + m * a.out_ld + a.out_choff + c, ov);
    }
}

// one thread per (segment, channel): three passes over the P positions (max, sum of exp + weighted mean, weighted
// variance); a warp reads 32 consecutive channels of a position, so every pass is coalesced
template <typename TL, typename TX>
__global__ void __launch_bounds__(128)
asp_pool_kernel(const AspPoolArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const long long b = blockIdx.y;
    if (c >= a.C) return;
    const TL *l = static_cast<const TL *>(a.logits) + b * a.P * (long long)a.l_ld + a.l_choff + c;
    const TX *x = static_cast<const TX *>(a.x) + b * a.P * (long long)a.x_ld + a.x_choff + c;
    float mx = -INFINITY;
    for (int p = 0; p < a.P; ++p) mx = fmaxf(mx, to_f32(l[(long long)p * a.l_ld]));
    float den = 0.f, num = 0.f;
    for (int p = 0; p < a.P; ++p) {
        const float e = expf(to_f32(l[(long long)p * a.l_ld]) - mx);
        den += e;
        num = fmaf(e, to_f32(x[(long long)p * a.x_ld]), num);
    }
    const float inv = 1.f / den, mean = num * inv;
    float var = 0.f;
    for (int p = 0; p < a.P; ++p) {
        const float e = expf(to_f32(l[(long long)p * a.l_ld]) - mx) * inv;
        const float d = to_f32(x[(long long)p * a.x_ld]) - mean;
        var = fmaf(e * d, d, var);
    }
    float *y = a.out + b * 2ll * a.C;
    y[c] = mean;
    y[a.C + c] = sqrtf(fmaxf(var, a.var_floor));
}

}  // namespace

int launch_se_scale(const SeScaleArgs &a, int dtype, int res_dtype, int out_dtype, cudaStream_t s) {
    if (a.C % 4 != 0 || a.x_ld % 4 || a.x_choff % 4 || a.out_ld % 4 || a.out_choff % 4 || (a.res != nullptr && (a.res_ld % 4 || a.res_choff % 4))) {
        set_error("se_scale: channel counts and offsets must be multiples of 4");
        return SPK_ERR_UNSUPPORTED;
    }
    const long long work = a.B * a.P * (a.C / 4);
    if (work == 0) return SPK_OK;
    const int g = (int)std::min<long long>((work + 255) / 256, 148ll * 32);
    if (a.res == nullptr) res_dtype = dtype;
    if (dtype == SPK_DT_F32 && res_dtype == SPK_DT_F32 && out_dtype == SPK_DT_F32) se_scale_kernel<float, float, float><<<g, 256, 0, s>>>(a);
    else if (dtype == SPK_DT_BF16 && res_dtype == SPK_DT_BF16 && out_dtype == SPK_DT_BF16) se_scale_kernel<bf16, bf16, bf16><<<g, 256, 0, s>>>(a);
    else {
        set_error("se_scale: unsupported dtype combination %d/%d/%d", dtype, res_dtype, out_dtype);
        return SPK_ERR_UNSUPPORTED;
    }
    return check_launch("se_scale_kernel");
}

int launch_asp_pool(const AspPoolArgs &a, int l_dtype, int x_dtype, cudaStream_t s) {
    if (a.B == 0) return SPK_OK;
    if (a.B > 65535) {
        set_error("asp_pool: batch too large for one launch (%lld)", a.B);
        return SPK_ERR_UNSUPPORTED;
    }
    if (asp_pool_online_supported(a)) return launch_asp_pool_online(a, l_dtype, x_dtype, s);
    dim3 grid((a.C + 127) / 128, (unsigned)a.B);
    if (l_dtype == SPK_DT_F32 && x_dtype == SPK_DT_F32) asp_pool_kernel<float, float><<<grid, 128, 0, s>>>(a);
    else if (l_dtype == SPK_DT_BF16 && x_dtype == SPK_DT_BF16) asp_pool_kernel<bf16, bf16><<<grid, 128, 0, s>>>(a);
    else if (l_dtype == SPK_DT_F32 && x_dtype == SPK_DT_BF16) asp_pool_kernel<float, bf16><<<grid, 128, 0, s>>>(a);
    else {
        set_error("asp_pool: unsupported dtype combination %d/%d", l_dtype, x_dtype);
        return SPK_ERR_UNSUPPORTED;
    }
    return check_launch("asp_pool_kernel");
}

}  // namespace spk

// ---- per-recording mean of chunk embeddings (speakerlab/bin/infer_sv_batch.py:313-315: embeddings[pos[i]:pos[i+1]].mean(0)).
// One CTA per recording, thread = embedding column: coalesced rows, sequential (fixed-order) accumulation.
namespace spk {
namespace {
__global__ void __launch_bounds__(256)
segment_mean_kernel(const float *__restrict__ E, int D, const int *__restrict__ pos, float *__restrict__ out) {
    const int w = blockIdx.x, lo = pos[w], hi = pos[w + 1];
    const float inv = hi > lo ? 1.0f / (float)(hi - lo) : 0.f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float s = 0.f;
        for (int r = lo; r < hi; ++r) s += E[(size_t)r * D + c];
        out[(size_t)w * D + c] = s * inv;
    }
}
}  // namespace
}  // namespace spk

extern "C" int spk_segment_mean(const float *E, int64_t D, const int32_t *pos, int64_t n_wav, float *out, void *stream) {
    using namespace spk;
    SPK_REQUIRE(n_wav >= 0 && D > 0, "bad sizes");
    if (n_wav == 0) return SPK_OK;
    SPK_REQUIRE(E != nullptr && pos != nullptr && out != nullptr, "null buffer");
    SPK_REQUIRE(n_wav < (1ll << 31), "too many recordings");
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    segment_mean_kernel<<<(unsigned)n_wav, 256, 0, static_cast<cudaStream_t>(stream)>>>(E, (int)D, pos, out);
    return check_launch("segment_mean_kernel");
}
