// Kaldi-compatible batched fbank + utterance CMN for sm_100a.
//
// Replaces FBank.__call__ (speakerlab/process/processor.py:143-158), i.e.
// torchaudio.compliance.kaldi.fbank (kaldi.py:514-645) with dither=0, snip_edges, remove_dc,
// pre-emphasis 0.97, povey window, 512-point power spectrum, triangular mel bank, log floor
// at FLT_EPSILON, followed by per-utterance mean normalisation.
//
// Mapping: one CTA per utterance (fused CMN: the whole [m, n_mels] log-mel tile lives in shared
// memory, so the output is written exactly once) or per frame range (long utterances, CMN as
// a second pass).  Inside a CTA each HALF-WARP owns one frame at a time:
//   * coalesced float2 loads of the 400 samples (frames overlap 2.5x; re-reads hit L1),
//   * DC removal / pre-emphasis / window in registers (neighbour sample via shuffles),
//   * 512-point real FFT as a 256-point complex FFT, 16 x 16 Cooley-Tukey: radix-16 in
//     registers, one 16x16 transpose through padded shared memory, radix-16 again,
//   * real-FFT untangle with the conjugate partner fetched by shuffle,
//   * power spectrum staged in shared memory, sparse triangular mel filters (<= 2 filters per
//     bin, 501 non-zeros for 80 bins) gathered per lane, log.
// HBM traffic is the algorithmic minimum: 4*n_samples in, 4*m*n_mels out per utterance.
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace spk {
namespace {

constexpr int kFrameLen = 400;
constexpr int kShift = 160;
constexpr int kBins = 256;
constexpr int kMaxMels = 128;
constexpr int kMaxNnz = 4096;
constexpr float kEps = 1.1920928955078125e-07f;
constexpr float kPreemph = 0.97f;

struct FbankTables {
    float window[kFrameLen];
    float2 tw256[256];   // exp(-2 pi i k / 256)
    float2 tw512[16];    // exp(-2 pi i k / 512), k < 16
    float2 tw32[16];     // exp(-2 pi i k / 32)
    int mel_start[kMaxMels];
    int mel_count[kMaxMels];
    int mel_woff[kMaxMels];
    float mel_w[kMaxNnz];
    int n_mels;
    int nnz;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// forward 4-point DFT in place
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y);
    float2 t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
    float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y);
    float2 t3 = make_float2(a1.x - a3.x, a1.y - a3.y);
    a0 = make_float2(t0.x + t2.x, t0.y + t2.y);
    a2 = make_float2(t0.x - t2.x, t0.y - t2.y);
    a1 = make_float2(t1.x + t3.y, t1.y - t3.x);
    a3 = make_float2(t1.x - t3.y, t1.y + t3.x);
}

// forward 16-point DFT in registers.  Input natural order v[n]; output X[k] is left at
// v[nat16(k)] with nat16(k) = 4*(k&3) + (k>>2).
__host__ __device__ constexpr int nat16(int k) { return 4 * (k & 3) + (k >> 2); }

__device__ __forceinline__ void fft16(float2 (&v)[16]) {
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; ++b) fft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // after this v[4c+b] = sum_a W4^{ac} x[4a+b]; twiddle by W16^{bc}
    v[4 * 1 + 1] = cmul(v[4 * 1 + 1], make_float2(c1, -s1));    // W^1
    v[4 * 1 + 2] = cmul(v[4 * 1 + 2], make_float2(r2, -r2));    // W^2
    v[4 * 1 + 3] = cmul(v[4 * 1 + 3], make_float2(s1, -c1));    // W^3
    v[4 * 2 + 1] = cmul(v[4 * 2 + 1], make_float2(r2, -r2));    // W^2
    v[4 * 2 + 2] = make_float2(v[4 * 2 + 2].y, -v[4 * 2 + 2].x); // W^4 = -i
    v[4 * 2 + 3] = cmul(v[4 * 2 + 3], make_float2(-r2, -r2));   // W^6
    v[4 * 3 + 1] = cmul(v[4 * 3 + 1], make_float2(s1, -c1));    // W^3
    v[4 * 3 + 2] = cmul(v[4 * 3 + 2], make_float2(-r2, -r2));   // W^6
    v[4 * 3 + 3] = cmul(v[4 * 3 + 3], make_float2(-c1, s1));    // W^9
#pragma unroll
    for (int c = 0; c < 4; ++c) fft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    // X[c + 4d] = v[4c + d]
}

constexpr int kXposePitch = 17;                       // float2 elements, conflict-free columns
constexpr int kWarpScratchFloats = 2 * 16 * kXposePitch * 2;   // 16x16 transpose per half-warp; the power
//                                                              spectrum (2 x 256 floats) aliases it once dead

// FUSED: CTA owns all frames of utterance blockIdx.x and applies CMN from shared memory.
// !FUSED: CTA owns frames [blockIdx.y*frames_per_cta, ...) and writes raw log-mel.
template <bool FUSED>
__global__ void __launch_bounds__(256)
fbank_kernel(const float *__restrict__ wav, int64_t n_samples, int64_t wav_stride,
             float *__restrict__ out, int m, int n_mels, int mean_nor, int frames_per_cta,
             const FbankTables *__restrict__ tab) {
    extern __shared__ __align__(16) float smem_raw[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw = lane >> 4, l = lane & 15;
    const int nnz = tab->nnz;

    float *p = smem_raw;
    float *out_s = p;
    if (FUSED) p += (size_t)m * n_mels;
    float *scratch = p + (size_t)warp * kWarpScratchFloats;
    p += (size_t)nwarps * kWarpScratchFloats;
    float *mel_w = p;
    p += (nnz + 3) & ~3;
    int *mel_start = reinterpret_cast<int *>(p);
    p += kMaxMels;
    int *mel_count = reinterpret_cast<int *>(p);
    p += kMaxMels;
    int *mel_woff = reinterpret_cast<int *>(p);
    p += kMaxMels;
    float *colmean = p;

    for (int i = threadIdx.x; i < nnz; i += blockDim.x) mel_w[i] = tab->mel_w[i];
    for (int i = threadIdx.x; i < n_mels; i += blockDim.x) {
        mel_start[i] = tab->mel_start[i];
        mel_count[i] = tab->mel_count[i];
        mel_woff[i] = tab->mel_woff[i];
    }
    __syncthreads();

    // per-lane constants
    float2 twl[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) twl[k] = tab->tw256[(l * k) & 255];
    const float2 tw5 = tab->tw512[l];

    float2 *xpose = reinterpret_cast<float2 *>(scratch) + hw * 16 * kXposePitch;
    float *pw = scratch + hw * kBins;   // aliases xpose (dead after the column reads)

    const int64_t b = blockIdx.x;
    const int f_begin = FUSED ? 0 : blockIdx.y * frames_per_cta;
    const int f_end = FUSED ? m : min(m, f_begin + frames_per_cta);
    const float *wrow = wav + b * wav_stride;
    float *orow = out + b * (int64_t)m * n_mels;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(wrow) & 7) == 0);

    for (int f0 = f_begin + warp * 2; f0 < f_end; f0 += nwarps * 2) {
        const int f = f0 + hw;
        const bool valid = f < f_end;
        const float *x = wrow + (int64_t)(valid ? f : f_end - 1) * kShift;

        // ---- load 400 samples as 200 complex values z[n] = x[2n] + i x[2n+1], n = 16*mm + l
        float2 z[16];
        float sum = 0.f;
#pragma unroll
        for (int mm = 0; mm < 16; ++mm) {
            const int i0 = 32 * mm + 2 * l;
            if (mm < 12 || (mm == 12 && l < 8)) {
                if (vec_ok) {
                    z[mm] = __ldg(reinterpret_cast<const float2 *>(x + i0));
                } else {
                    z[mm] = make_float2(__ldg(x + i0), __ldg(x + i0 + 1));
                }
                sum += z[mm].x + z[mm].y;
            } else {
                z[mm] = make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, 16);
        const float mean = sum * (1.0f / kFrameLen);

        // ---- DC removal, pre-emphasis (replicate at i = 0), povey window
#pragma unroll
        for (int mm = 0; mm < 13; ++mm) {
            z[mm].x -= mean;
            z[mm].y -= mean;
        }
        float carry = z[0].x;   // d[-1] := d[0]
#pragma unroll
        for (int mm = 0; mm < 13; ++mm) {
            const float up = __shfl_up_sync(0xffffffffu, z[mm].y, 1, 16);
            const float prev = (l == 0) ? carry : up;
            carry = __shfl_sync(0xffffffffu, z[mm].y, 15, 16);   // lane 15's imag feeds lane 0 next row
            const int i0 = 32 * mm + 2 * l;
            float2 w = make_float2(0.f, 0.f);
            if (mm < 12 || l < 8) w = __ldg(reinterpret_cast<const float2 *>(tab->window + i0));
            const float yr = z[mm].x - kPreemph * prev;
            const float yi = z[mm].y - kPreemph * z[mm].x;
            z[mm] = make_float2(yr * w.x, yi * w.y);
        }

        // ---- 256-point complex FFT = 16 (registers) x 16 (lanes)
        fft16(z);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) xpose[k1 * kXposePitch + l] = cmul(z[nat16(k1)], twl[k1]);
        __syncwarp();
        float2 u[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) u[j] = xpose[l * kXposePitch + j];
        __syncwarp();   // xpose is dead from here on; pw reuses its storage
        fft16(u);
        // lane l now holds Z[l + 16*k2] in u[nat16(k2)]

        // ---- real-FFT untangle + power spectrum
        const int partner = (16 - l) & 15;
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            const float2 zk = u[nat16(k2)];
            const float2 other = u[nat16(15 - k2)];
            float2 zp;
            zp.x = __shfl_sync(0xffffffffu, other.x, partner, 16);
            zp.y = __shfl_sync(0xffffffffu, other.y, partner, 16);
            if (l == 0) zp = u[nat16((16 - k2) & 15)];
            const float2 a = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
            const float2 bc = make_float2(0.5f * (zk.x - zp.x), 0.5f * (zk.y + zp.y));
            const float2 wk = cmul(tw5, tab->tw32[k2]);
            const float2 q = cmul(wk, bc);
            const float xr = a.x + q.y, xi = a.y - q.x;
            pw[l + 16 * k2] = xr * xr + xi * xi;
        }
        __syncwarp();

        // ---- sparse triangular mel filters + log
        for (int s = 0; s * 16 < n_mels; ++s) {
            const int fi = 16 * s + ((s & 1) ? 15 - l : l);
            if (fi < n_mels) {
                const int st = mel_start[fi], cnt = mel_count[fi];
                const float *w = mel_w + mel_woff[fi];
                float e = 0.f;
                for (int t = 0; t < cnt; ++t) e = fmaf(pw[st + t], w[t], e);
                const float v = logf(fmaxf(e, kEps));
                if (valid) {
                    if (FUSED) out_s[(size_t)f * n_mels + fi] = v;
                    else orow[(size_t)f * n_mels + fi] = v;
                }
            }
        }
        __syncwarp();
    }

    if (!FUSED) return;
    __syncthreads();
    // ---- utterance CMN (processor.py:156-157) from shared memory, then one coalesced store
    const int total = m * n_mels;
    if (mean_nor) {
        for (int c = threadIdx.x; c < n_mels; c += blockDim.x) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int fr = 0;
            for (; fr + 3 < m; fr += 4) {
                s0 += out_s[(size_t)fr * n_mels + c];
                s1 += out_s[(size_t)(fr + 1) * n_mels + c];
                s2 += out_s[(size_t)(fr + 2) * n_mels + c];
                s3 += out_s[(size_t)(fr + 3) * n_mels + c];
            }
            for (; fr < m; ++fr) s0 += out_s[(size_t)fr * n_mels + c];
            colmean[c] = ((s0 + s1) + (s2 + s3)) / (float)m;
        }
        __syncthreads();
    }
    const bool st_vec = ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) && (n_mels % 4 == 0);
    if (st_vec) {
        for (int i = threadIdx.x * 4; i < total; i += blockDim.x * 4) {
            float4 v = *reinterpret_cast<const float4 *>(out_s + i);
            if (mean_nor) {
                const int c = i % n_mels;
                v.x -= colmean[c]; v.y -= colmean[c + 1]; v.z -= colmean[c + 2]; v.w -= colmean[c + 3];
            }
            __stcs(reinterpret_cast<float4 *>(orow + i), v);
        }
    } else {
        for (int i = threadIdx.x; i < total; i += blockDim.x)
            orow[i] = out_s[i] - (mean_nor ? colmean[i % n_mels] : 0.f);
    }
}

// second pass for long utterances: out[b, :, c] -= mean over frames
__global__ void __launch_bounds__(256)
cmn_kernel(float *__restrict__ out, int m, int n_mels) {
    extern __shared__ float sh[];   // [rows_per_iter = blockDim.x / 32 ... ] partial sums
    float *orow = out + (int64_t)blockIdx.x * m * n_mels;
    float *colsum = sh;             // n_mels
    for (int c = threadIdx.x; c < n_mels; c += blockDim.x) colsum[c] = 0.f;
    __syncthreads();
    // each thread owns column (tid % n_mels) over a strided set of frames
    const int cols = n_mels;
    const int groups = blockDim.x / cols;
    if (groups > 0 && threadIdx.x < groups * cols) {
        const int c = threadIdx.x % cols, g = threadIdx.x / cols;
        float s = 0.f;
        for (int fr = g; fr < m; fr += groups) s += orow[(size_t)fr * n_mels + c];
        atomicAdd(&colsum[c], s);
    } else if (groups == 0) {
        for (int c = threadIdx.x; c < n_mels; c += blockDim.x) {
            float s = 0.f;
            for (int fr = 0; fr < m; ++fr) s += orow[(size_t)fr * n_mels + c];
            colsum[c] = s;
        }
    }
    __syncthreads();
    const int total = m * n_mels;
    for (int i = threadIdx.x; i < total; i += blockDim.x) orow[i] -= colsum[i % n_mels] / (float)m;
}

// ------------------------------------------------------------------ host tables
struct HostState {
    std::mutex mu;
    std::vector<float> window;            // 400
    std::vector<float> mel;               // [n_mels, 256] override (empty = built-in)
    int override_mels = 0;
    bool window_override = false;
    // device copies, one per (device, n_mels)
    struct Dev { int device; int n_mels; FbankTables *ptr; uint64_t gen; int nnz; };
    std::vector<Dev> devs;
    uint64_t gen = 1;
};
HostState &state() {
    static HostState s;
    return s;
}

void builtin_window(float *w) {
    for (int i = 0; i < kFrameLen; ++i) {
        double h = 0.5 - 0.5 * std::cos(2.0 * M_PI * i / (kFrameLen - 1));
        w[i] = (float)std::pow(h, 0.85);
    }
}

// kaldi.py:436-511, evaluated in double and rounded once
void builtin_mel(int n_mels, std::vector<float> &mel) {
    mel.assign((size_t)n_mels * kBins, 0.f);
    auto melscale = [](double f) { return 1127.0 * std::log(1.0 + f / 700.0); };
    const double lo = melscale(20.0), hi = melscale(8000.0);
    const double delta = (hi - lo) / (n_mels + 1);
    const double bw = 16000.0 / 512.0;
    for (int j = 0; j < n_mels; ++j) {
        const double left = lo + j * delta, center = lo + (j + 1) * delta, right = lo + (j + 2) * delta;
        for (int k = 0; k < kBins; ++k) {
            const double mk = melscale(bw * k);
            const double up = (mk - left) / (center - left), down = (right - mk) / (right - center);
            const double v = std::fmax(0.0, std::fmin(up, down));
            mel[(size_t)j * kBins + k] = (float)v;
        }
    }
}

int build_tables(int n_mels, FbankTables &t) {
    HostState &s = state();
    std::vector<float> mel;
    if (s.override_mels == n_mels && !s.mel.empty()) mel = s.mel;
    else builtin_mel(n_mels, mel);
    if (s.window_override) std::copy(s.window.begin(), s.window.end(), t.window);
    else builtin_window(t.window);
    for (int k = 0; k < 256; ++k) {
        double a = -2.0 * M_PI * k / 256.0;
        t.tw256[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    for (int k = 0; k < 16; ++k) {
        double a = -2.0 * M_PI * k / 512.0, c = -2.0 * M_PI * k / 32.0;
        t.tw512[k] = make_float2((float)std::cos(a), (float)std::sin(a));
        t.tw32[k] = make_float2((float)std::cos(c), (float)std::sin(c));
    }
    int nnz = 0;
    for (int j = 0; j < n_mels; ++j) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k)
            if (mel[(size_t)j * kBins + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        const int cnt = first < 0 ? 0 : last - first + 1;
        if (nnz + cnt > kMaxNnz) {
            set_error("mel bank too dense (%d non-zeros > %d)", nnz + cnt, kMaxNnz);
            return SPK_ERR_UNSUPPORTED;
        }
        t.mel_start[j] = first < 0 ? 0 : first;
        t.mel_count[j] = cnt;
        t.mel_woff[j] = nnz;
        for (int k = 0; k < cnt; ++k) t.mel_w[nnz + k] = mel[(size_t)j * kBins + first + k];
        nnz += cnt;
    }
    t.n_mels = n_mels;
    t.nnz = nnz;
    return SPK_OK;
}

int get_tables(int n_mels, const FbankTables **out, int *nnz_out) {
    HostState &s = state();
    std::lock_guard<std::mutex> lk(s.mu);
    int dev = 0;
    SPK_CUDA_OK(cudaGetDevice(&dev));
    for (auto &d : s.devs)
        if (d.device == dev && d.n_mels == n_mels && d.gen == s.gen) {
            *out = d.ptr;
            *nnz_out = d.nnz;
            return SPK_OK;
        }
    static FbankTables host_t;   // guarded by s.mu
    int rc = build_tables(n_mels, host_t);
    if (rc != SPK_OK) return rc;
    FbankTables *dptr = nullptr;
    SPK_CUDA_OK(cudaMalloc(&dptr, sizeof(FbankTables)));
    SPK_CUDA_OK(cudaMemcpy(dptr, &host_t, sizeof(FbankTables), cudaMemcpyHostToDevice));
    s.devs.push_back({dev, n_mels, dptr, s.gen, host_t.nnz});
    *out = dptr;
    *nnz_out = host_t.nnz;
    return SPK_OK;
}

size_t smem_bytes(bool fused, int m, int n_mels, int nwarps, int nnz) {
    size_t fl = 0;
    if (fused) fl += (size_t)m * n_mels;
    fl += (size_t)nwarps * kWarpScratchFloats;
    fl += (nnz + 3) & ~3;
    fl += 3 * kMaxMels;
    fl += kMaxMels;   // colmean
    return fl * sizeof(float);
}

}  // namespace
}  // namespace spk

using namespace spk;

extern "C" int64_t spk_fbank_num_frames(int64_t n_samples) {
    if (n_samples < kFrameLen) return 0;
    return 1 + (n_samples - kFrameLen) / kShift;
}

extern "C" int spk_fbank_set_tables(const float *window400, const float *mel_bank, int n_mels) {
    HostState &s = state();
    std::lock_guard<std::mutex> lk(s.mu);
    if (mel_bank != nullptr) {
        SPK_REQUIRE(n_mels >= 1 && n_mels <= kMaxMels, "n_mels %d out of range [1,%d]", n_mels, kMaxMels);
        s.mel.assign(mel_bank, mel_bank + (size_t)n_mels * kBins);
        s.override_mels = n_mels;
    } else {
        s.mel.clear();
        s.override_mels = 0;
    }
    if (window400 != nullptr) {
        s.window.assign(window400, window400 + kFrameLen);
        s.window_override = true;
    } else {
        s.window_override = false;
    }
    s.gen++;   // invalidate device copies (old ones are leaked on purpose: a few KB, rare)
    return SPK_OK;
}

extern "C" int spk_fbank_f32(const float *wav, int64_t B, int64_t n_samples, int64_t wav_stride,
                             float *out, int n_mels, int mean_nor, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SPK_REQUIRE(B >= 0, "negative batch");
    SPK_REQUIRE(B == 0 || (wav != nullptr && out != nullptr), "null buffer");
    // kaldi.py:142: assert 2 <= window_size <= len(waveform)
    SPK_REQUIRE(n_samples >= kFrameLen, "choose a window size %d that is [2, %lld]", kFrameLen,
                (long long)n_samples);
    SPK_REQUIRE(wav_stride >= n_samples, "wav_stride %lld < n_samples %lld", (long long)wav_stride,
                (long long)n_samples);
    SPK_REQUIRE(n_mels > 3 && n_mels <= kMaxMels, "n_mels %d out of range (3,%d]", n_mels, kMaxMels);
    if (B == 0) return SPK_OK;
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    const FbankTables *tab = nullptr;
    int nnz = 0;
    rc = get_tables(n_mels, &tab, &nnz);
    if (rc != SPK_OK) return rc;

    const int64_t m64 = spk_fbank_num_frames(n_samples);
    SPK_REQUIRE(m64 < (1ll << 30), "too many frames");
    const int m = (int)m64;
    const int threads = 256, nwarps = threads / 32;
    const size_t fused_bytes = smem_bytes(true, m, n_mels, nwarps, nnz);
    static const size_t kFusedLimit = 100 * 1024;   // 2 CTAs / SM
    if (fused_bytes <= kFusedLimit || (fused_bytes <= 200 * 1024 && B >= 2 * sm_count())) {
        static std::once_flag once;
        static cudaError_t attr_err = cudaSuccess;
        std::call_once(once, [] {
            attr_err = cudaFuncSetAttribute(fbank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            220 * 1024);
        });
        SPK_CUDA_OK(attr_err);
        // grid.x is limited to 2^31-1; B beyond that is not a realistic batch
        SPK_REQUIRE(B < (1ll << 31), "batch too large");
        fbank_kernel<true><<<(unsigned)B, threads, fused_bytes, stream>>>(wav, n_samples, wav_stride, out, m,
                                                                           n_mels, mean_nor, m, tab);
        return check_launch("fbank_kernel<fused>");
    }
    // long utterances: frame-range CTAs + second-pass CMN
    const int frames_per_cta = 64;
    const int chunks = (m + frames_per_cta - 1) / frames_per_cta;
    SPK_REQUIRE(chunks <= 65535, "utterance too long (%d frames)", m);
    dim3 grid((unsigned)B, (unsigned)chunks);
    fbank_kernel<false><<<grid, threads, smem_bytes(false, m, n_mels, nwarps, nnz), stream>>>(
        wav, n_samples, wav_stride, out, m, n_mels, mean_nor, frames_per_cta, tab);
    rc = check_launch("fbank_kernel<ranges>");
    if (rc != SPK_OK) return rc;
    if (mean_nor) {
        cmn_kernel<<<(unsigned)B, 256, kMaxMels * sizeof(float), stream>>>(out, m, n_mels);
        rc = check_launch("cmn_kernel");
    }
    return rc;
}

extern "C" int spk_fbank_host_f32(const float *wav, int64_t B, int64_t n_samples, int64_t wav_stride,
                                  float *out, int n_mels, int mean_nor) {
    SPK_REQUIRE(B == 0 || (wav != nullptr && out != nullptr), "null buffer");
    SPK_REQUIRE(n_samples >= kFrameLen, "choose a window size %d that is [2, %lld]", kFrameLen,
                (long long)n_samples);
    if (B == 0) return SPK_OK;
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    const int64_t m = spk_fbank_num_frames(n_samples);
    float *dwav = nullptr, *dout = nullptr;
    const size_t in_b = (size_t)B * n_samples * sizeof(float), out_b = (size_t)B * m * n_mels * sizeof(float);
    SPK_CUDA_OK(cudaMalloc(&dwav, in_b));
    cudaError_t e = cudaMalloc(&dout, out_b);
    if (e != cudaSuccess) {
        cudaFree(dwav);
        set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
        return SPK_ERR_CUDA;
    }
    cudaStream_t s = nullptr;
    e = cudaMemcpy2DAsync(dwav, n_samples * sizeof(float), wav, wav_stride * sizeof(float),
                          n_samples * sizeof(float), B, cudaMemcpyHostToDevice, s);
    rc = SPK_OK;
    if (e == cudaSuccess) rc = spk_fbank_f32(dwav, B, n_samples, n_samples, dout, n_mels, mean_nor, s);
    if (e == cudaSuccess && rc == SPK_OK) e = cudaMemcpyAsync(out, dout, out_b, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && rc == SPK_OK) e = cudaStreamSynchronize(s);
    cudaFree(dwav);
    cudaFree(dout);
    if (rc != SPK_OK) return rc;
    if (e != cudaSuccess) {
        set_error("fbank host path failed: %s", cudaGetErrorString(e));
        return SPK_ERR_CUDA;
    }
    return SPK_OK;
}
