// Kaldi-compatible batched fbank + utterance CMN for sm_100a.
//
// Replaces FBank.__call__ (speakerlab/process/processor.py:143-158), i.e.
// torchaudio.compliance.kaldi.fbank (kaldi.py:514-645) with dither=0, snip_edges, remove_dc,
// pre-emphasis 0.97, povey window, 512-point power spectrum, triangular mel bank, log floor
// at FLT_EPSILON, followed by per-utterance mean normalisation.  Input rows are float32 in
// [-1,1] scale or int16 PCM (scaled by 1/32768 on load, speakerlab/utils/fileio.py:115-117), either a
// [B, n] matrix or windows (start, length) of ONE resident recording with the reference's
// circle_pad for short tail windows (infer_diarization.py:621-627, utils.py:232-238).
//
// Arithmetic.  The reference transforms v[n] = w[n] (d[n] - 0.97 d[n-1]); for the low bins that signal is ~30x
// below the raw spectrum level, so a float32 FFT of v leaves them with 1e-3-level log errors (so does the
// reference's own float32 path).  With w[0] = w[399] = 0 the transform splits exactly into
//       X[k] = (1 - 0.97 e^{-2 pi i k/512}) B[k] + C[k],   B = DFT(d[m] w[m+1]),  C = DFT(d[m] (w[m] - w[m+1]))
// where B has no cancellation and C is ~1/64 of B's level.  One COMPLEX 512-point FFT of
// z = d w+ + i 64 d dw gives both (B and C are its conjugate-symmetric / antisymmetric parts), so every bin
// comes out with float32 RELATIVE accuracy.  Cells whose mel energy is still > 27 dB below the white-noise
// expectation of the frame (where no float32 method reaches 1e-4 in the log) are recomputed in float64 by a
// direct DFT (a few cells per million on noise-like input, capped per round).
//
// Mapping.  CTA = 256 threads = 16 half-warps; a round is 16 frames, one per half-warp:
//   * lane c loads samples 16a + c (a < 25), removes the frame mean, multiplies by the (w+, 64 dw) table -> z,
//   * 32-point DFT over a in registers (radix-2 + two radix-16), twiddle, 32x16 transpose through padded
//     shared memory, lane c takes rows c and 32-c and runs two 16-point DFTs: it then owns Z[k] AND Z[512-k]
//     for its 16 bins, so the split into B and C needs no cross-lane traffic; all butterflies are packed
//     f32x2 instructions (one complex add per instruction),
//   * power spectrum -> shared memory [bin][frame]; after a CTA barrier thread (frame, filter group) runs the
//     sparse triangular filters (<= 2 filters per bin; weights are warp-uniform broadcasts, powers
//     conflict-free), log, and the 16 x n_mels tile leaves through one coalesced store.
//   * CMN: column sums ride along in registers; a CTA that owns the whole utterance subtracts the mean in a
//     second pass over its own (L2-resident) rows, otherwise a small second kernel does.
// HBM traffic is the algorithmic minimum (4 n in, 4 m n_mels out per utterance); the kernel is bound by the
// fp32 pipe (about 640 packed instructions per lane per frame).
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
constexpr int kThreads = 256;
constexpr int kFramesPerRound = 16;            // one per half-warp
constexpr int kGroups = 16;                    // filter groups in the mel phase (threads = frames x groups)
constexpr int kMaxSlots = kMaxMels / kGroups;  // filters per group
constexpr int kMaxW = 8192;                    // padded mel weights
constexpr int kXpPitch = 17;                   // float2 per transpose row (16 + 1)
constexpr int kXpFloat2 = 32 * kXpPitch;       // per half-warp
constexpr int kPwPitch = 17;                   // floats per bin row (16 frames + 1)
constexpr int kPwRows = kBins + 32;            // zero rows behind the spectrum absorb padded filter taps
constexpr int kRepairMax = 256;                // list capacity; the active cap is a run-time knob (spk_fbank_set_repair)
constexpr float kEps = 1.1920928955078125e-07f;
constexpr double kPreemph = 0.97;
constexpr float kScaleC = 64.f;                // C is carried as 64 C in the imaginary input

struct FbankTables {
    float2 T[kFrameLen];                 // (w[n+1], 64 (w[n] - w[n+1]))
    float2 tw[32 * 16];                  // tw[k1 * 16 + c] = exp(-2 pi i c k1 / 512)
    float2 Hh[kBins];                    // (1 - 0.97 exp(-2 pi i k / 512)) / 2
    float melw[kMaxW];                   // per (group, slot) weight runs, zero padded to the warp's trip count
    int slot_filter[kGroups * kMaxSlots];   // filter index or -1
    int slot_start[kGroups * kMaxSlots];    // first bin
    int slot_woff[kGroups * kMaxSlots];
    int slot_trip[(kGroups / 2) * kMaxSlots];   // taps run by the warp that owns groups 2w, 2w+1
    float gE[kMaxMels];                  // sum_k mel[j][k] |H[k]|^2: expected energy per unit frame energy
    int mel_start[kMaxMels], mel_count[kMaxMels], mel_woff[kMaxMels];   // exact runs (float64 repair)
    float mel_w[4096];
    double win64[kFrameLen];
    int n_mels, n_slots, n_w;
};

// ---- packed complex helpers (float2 = one 64-bit register pair; add/mul/fma are single f32x2 instructions)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }   // exact a - b
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    // (a.x b.x - a.y b.y, a.x b.y + a.y b.x) = fma((a.y, a.y), (-b.y, b.x), (a.x, a.x) * b)
    return __ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), __fmul2_rn(make_float2(a.x, a.x), b));
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }     // a * (-i)

// forward 4-point DFT in place
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

// forward 16-point DFT in registers.  Input natural order v[n]; output X[k] is left at v[nat16(k)].
__host__ __device__ constexpr int nat16(int k) { return 4 * (k & 3) + (k >> 2); }

__device__ __forceinline__ void fft16(float2 (&v)[16]) {
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; ++b) fft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // v[4c+b] = sum_a W4^{ac} x[4a+b]; twiddle by W16^{bc}
    v[4 * 1 + 1] = cmul(v[4 * 1 + 1], make_float2(c1, -s1));
    v[4 * 1 + 2] = cmul(v[4 * 1 + 2], make_float2(r2, -r2));
    v[4 * 1 + 3] = cmul(v[4 * 1 + 3], make_float2(s1, -c1));
    v[4 * 2 + 1] = cmul(v[4 * 2 + 1], make_float2(r2, -r2));
    v[4 * 2 + 2] = mul_mi(v[4 * 2 + 2]);
    v[4 * 2 + 3] = cmul(v[4 * 2 + 3], make_float2(-r2, -r2));
    v[4 * 3 + 1] = cmul(v[4 * 3 + 1], make_float2(s1, -c1));
    v[4 * 3 + 2] = cmul(v[4 * 3 + 2], make_float2(-r2, -r2));
    v[4 * 3 + 3] = cmul(v[4 * 3 + 3], make_float2(-c1, s1));
#pragma unroll
    for (int c = 0; c < 4; ++c) fft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// exp(-2 pi i a / 32), a < 16
__device__ __forceinline__ float2 w32(int a) {
    constexpr float cs[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                              0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                              0.f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                              -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
    constexpr float sn[16] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                              0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                              1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                              0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
    return make_float2(cs[a], -sn[a]);
}

template <typename T> __device__ __forceinline__ float load_sample(const T *p);
template <> __device__ __forceinline__ float load_sample<float>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_sample<int16_t>(const int16_t *p) {
    return __fmul_rn((float)__ldg(p), 1.0f / 32768.0f);       // exact
}

struct FbankArgs {
    const void *wav;            // float or int16 samples
    const int64_t *starts;      // optional [B]: origin of row b inside wav (window mode); else b * stride
    const int32_t *lens;        // optional [B]: period of row b: sample i reads origin + (phase + i) % period
    const int32_t *phases;      // optional [B]: phase of row b inside its period (0 when null)
    int64_t stride;             // row pitch in samples (matrix mode)
    int64_t n_samples;          // logical samples per row
    float *out;                 // [B, m, n_mels]
    int m, n_mels, mean_nor, frames_per_cta, fused;
    float repair_theta;         // repair when e < theta * E[e | white noise with the frame's energy]
    int repair_cap;             // float64 repairs per 16-frame round (more = pathological input, rest stays fp32)
    const FbankTables *tab;
};

__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    return s;
}

template <typename SampleT, bool WRAP>
__global__ void __launch_bounds__(kThreads, 2)
fbank_kernel(const FbankArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FbankTables *__restrict__ tab = A.tab;
    const int n_mels = A.n_mels, n_slots = tab->n_slots, n_w = tab->n_w;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int hw = tid >> 4, c = tid & 15;          // half-warp = frame slot, lane inside it

    // ---- shared-memory carve-up
    float2 *sT = reinterpret_cast<float2 *>(smem_raw);                 // 400 (+ pad to 416)
    float2 *sTw = sT + 416;                                            // 512
    float2 *sHh = sTw + 512;                                           // 256
    float2 *sXp = sHh + kBins;                                         // 16 x 32 x 17
    float *sPw = reinterpret_cast<float *>(sXp + kFramesPerRound * kXpFloat2);   // kPwRows x 17
    float *sStage = sPw + kPwRows * kPwPitch;                          // 16 x (n_mels + 1)
    float *sMelW = sStage + kFramesPerRound * (n_mels + 1);            // n_w (rounded up to 4)
    int *sSlot = reinterpret_cast<int *>(sMelW + ((n_w + 3) & ~3));    // filter/start/woff [16*8] x 3, trip [8*8]
    float *sEf = reinterpret_cast<float *>(sSlot + 3 * kGroups * kMaxSlots + (kGroups / 2) * kMaxSlots);   // 16
    float *sGE = sEf + kFramesPerRound;                                // n_mels: repair threshold per unit frame energy
    float *sColMean = sGE + kMaxMels;                                  // n_mels
    int *sRepair = reinterpret_cast<int *>(sColMean + kMaxMels);       // count + cap entries
    double *sRed = reinterpret_cast<double *>(sRepair + 2 + kRepairMax);   // 8 (8-byte aligned by construction)

    for (int i = tid; i < kFrameLen; i += kThreads) sT[i] = tab->T[i];
    for (int i = tid; i < 512; i += kThreads) sTw[i] = tab->tw[i];
    for (int i = tid; i < kBins; i += kThreads) sHh[i] = tab->Hh[i];
    for (int i = tid; i < n_w; i += kThreads) sMelW[i] = tab->melw[i];
    for (int i = tid; i < kGroups * kMaxSlots; i += kThreads) {
        sSlot[i] = tab->slot_filter[i];
        sSlot[kGroups * kMaxSlots + i] = tab->slot_start[i];
        sSlot[2 * kGroups * kMaxSlots + i] = tab->slot_woff[i];
    }
    for (int i = tid; i < (kGroups / 2) * kMaxSlots; i += kThreads) sSlot[3 * kGroups * kMaxSlots + i] = tab->slot_trip[i];
    for (int i = tid; i < kPwRows * kPwPitch; i += kThreads) sPw[i] = 0.f;
    for (int i = tid; i < kMaxMels; i += kThreads) sGE[i] = i < n_mels ? A.repair_theta * tab->gE[i] : 0.f;
    if (tid == 0) sRepair[0] = 0;
    __syncthreads();

    const int64_t b = blockIdx.x;
    const int f_begin = A.fused ? 0 : blockIdx.y * A.frames_per_cta;
    const int f_end = A.fused ? A.m : min(A.m, f_begin + A.frames_per_cta);
    const int64_t row0 = A.starts ? A.starts[b] : b * A.stride;
    const int row_len = A.lens ? A.lens[b] : (int)(A.n_samples < 0x7fffffffll ? A.n_samples : 0x7fffffffll);
    const int phase = (WRAP && A.phases) ? A.phases[b] : 0;
    const bool wrap = WRAP && (int64_t)phase + A.n_samples > (int64_t)row_len;   // circle_pad (utils.py:232-238)
    const SampleT *wrow = static_cast<const SampleT *>(A.wav) + row0;
    float *orow = A.out + b * (int64_t)A.m * n_mels;
    const int stage_pitch = n_mels + 1;

    float2 *xp = sXp + hw * kXpFloat2;
    const int rA = c, rB = (c == 0) ? 16 : 32 - c;
    // mel-phase role
    const int mf = tid & 15, mg = tid >> 4;
    float colsum[kMaxSlots];
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) colsum[s] = 0.f;

    for (int f0 = f_begin; f0 < f_end; f0 += kFramesPerRound) {
        // =============================================================== FFT phase: frame f0 + hw
        {
            const int f = f0 + hw;
            const bool valid = f < f_end;
            const int base = phase + (valid ? f : f_end - 1) * kShift;   // sample index inside the row's period (< 2^31)
            float x[25];
            float sum = 0.f;
#pragma unroll
            for (int a = 0; a < 25; ++a) {
                int i = base + 16 * a + c;
                if (wrap) i = (int)((unsigned)i % (unsigned)row_len);
                x[a] = load_sample<SampleT>(wrow + i);
                sum = __fadd_rn(sum, x[a]);       // explicit roundings: float and int16 inputs must agree bit for bit
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, 16);
            const float mean = __fmul_rn(sum, 1.0f / kFrameLen);       // rounded like the reference's x.mean()

            float2 E[16], O[16];
            float ef = 0.f;
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                const float d0 = __fsub_rn(x[a], mean);
                const float2 lo = __fmul2_rn(make_float2(d0, d0), sT[16 * a + c]);
                ef = fmaf(lo.x, lo.x, ef);
                if (a + 16 < 25) {
                    const float d1 = __fsub_rn(x[a + 16], mean);
                    const float2 hi = __fmul2_rn(make_float2(d1, d1), sT[16 * (a + 16) + c]);
                    ef = fmaf(hi.x, hi.x, ef);
                    E[a] = cadd(lo, hi);
                    O[a] = csub(lo, hi);
                } else {
                    E[a] = lo;
                    O[a] = lo;
                }
                if (a == 8) O[a] = mul_mi(O[a]);
                else if (a > 0) O[a] = cmul(O[a], w32(a));
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ef += __shfl_xor_sync(0xffffffffu, ef, o, 16);
            if (c == 0) sEf[hw] = ef;
            fft16(E);       // Y[2k'] at E[nat16(k')]
            fft16(O);       // Y[2k'+1] at O[nat16(k')]
#pragma unroll
            for (int kp = 0; kp < 16; ++kp) {
                float2 e = E[nat16(kp)], o = O[nat16(kp)];
                if (kp > 0) e = cmul(e, sTw[(2 * kp) * 16 + c]);
                o = cmul(o, sTw[(2 * kp + 1) * 16 + c]);
                xp[(2 * kp) * kXpPitch + c] = e;
                xp[(2 * kp + 1) * kXpPitch + c] = o;
            }
            __syncwarp();
            float2 P[16], Q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                P[j] = xp[rA * kXpPitch + j];
                Q[j] = xp[rB * kXpPitch + j];
            }
            __syncwarp();
            fft16(P);       // Z[rA + 32 k2] at P[nat16(k2)]
            fft16(Q);       // Z[rB + 32 k2] at Q[nat16(k2)]
            const float cs = 0.5f / kScaleC;
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const int bin = (which ? rB : rA) + 32 * k2;
                    const float2 a = which ? Q[nat16(k2)] : P[nat16(k2)];
                    float2 p;
                    if (which == 0) {
                        const float2 p0 = P[nat16((16 - k2) & 15)], p1 = Q[nat16(15 - k2)];
                        p = (c == 0) ? p0 : p1;
                    } else {
                        const float2 p0 = Q[nat16(15 - k2)], p1 = P[nat16(15 - k2)];
                        p = (c == 0) ? p0 : p1;
                    }
                    const float2 S = cadd(a, p), D = csub(a, p);
                    const float2 h = sHh[bin];
                    // X = Hh * (S.x, D.y) + cs * (S.y, -D.x)
                    float2 X = __fmul2_rn(make_float2(cs, -cs), make_float2(S.y, D.x));
                    X = __ffma2_rn(make_float2(S.x, S.x), h, X);
                    X = __ffma2_rn(make_float2(D.y, D.y), make_float2(-h.y, h.x), X);
                    sPw[bin * kPwPitch + hw] = fmaf(X.x, X.x, X.y * X.y);
                }
            }
        }
        __syncthreads();
        // =============================================================== mel phase: thread (frame mf, group mg)
        {
            const int f = f0 + mf;
            const bool valid = f < f_end;
            const float ef = sEf[mf];
#pragma unroll
            for (int s = 0; s < kMaxSlots; ++s) {
                if (s >= n_slots) break;
                const int q = mg * kMaxSlots + s;
                const int j = sSlot[q], st = sSlot[kGroups * kMaxSlots + q], wo = sSlot[2 * kGroups * kMaxSlots + q];
                const int trip = sSlot[3 * kGroups * kMaxSlots + warp * kMaxSlots + s];
                const float4 *w4 = reinterpret_cast<const float4 *>(sMelW + wo);      // runs are zero padded to 4 taps
                const float *pw = sPw + st * kPwPitch + mf;
                float e0 = 0.f, e1 = 0.f;
                for (int t = 0; t < trip; t += 4) {
                    const float4 w = w4[t >> 2];
                    e0 = fmaf(w.x, pw[t * kPwPitch], e0);
                    e1 = fmaf(w.y, pw[(t + 1) * kPwPitch], e1);
                    e0 = fmaf(w.z, pw[(t + 2) * kPwPitch], e0);
                    e1 = fmaf(w.w, pw[(t + 3) * kPwPitch], e1);
                }
                const float e = e0 + e1;
                if (j >= 0) {
                    const float v = logf(fmaxf(e, kEps));
                    sStage[mf * stage_pitch + j] = v;
                    if (valid) {
                        if (e < sGE[j] * ef) {
                            const int slot = atomicAdd(&sRepair[0], 1);
                            if (slot < A.repair_cap) sRepair[1 + slot] = (mf << 8) | j;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // =============================================================== float64 repair of flagged cells (rare)
        const int n_rep = min(sRepair[0], A.repair_cap);
        for (int r = 0; r < n_rep; ++r) {
            const int code = sRepair[1 + r];
            const int rf = code >> 8, j = code & 255;
            const int base = phase + (f0 + rf) * kShift;
            double xs[2], s = 0.0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = tid + h * kThreads;
                xs[h] = 0.0;
                if (n < kFrameLen) {
                    int i = base + n;
                    if (wrap) i = (int)((unsigned)i % (unsigned)row_len);
                    xs[h] = (double)load_sample<SampleT>(wrow + i);
                    s += xs[h];
                }
            }
            const double mean = block_sum(s, sRed) / kFrameLen;
            double vs[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = tid + h * kThreads;
                vs[h] = 0.0;
                if (n < kFrameLen) {
                    int i = base + max(n - 1, 0);
                    if (wrap) i = (int)((unsigned)i % (unsigned)row_len);
                    const double prev = (double)load_sample<SampleT>(wrow + i) - mean;
                    vs[h] = ((xs[h] - mean) - kPreemph * prev) * tab->win64[n];
                }
            }
            const int st = tab->mel_start[j], cnt = tab->mel_count[j];
            const float *mw = tab->mel_w + tab->mel_woff[j];
            double e = 0.0;
            for (int t = 0; t < cnt; ++t) {
                const int k = st + t;
                double re = 0.0, im = 0.0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = tid + h * kThreads;
                    if (n < kFrameLen) {
                        double sn, cn;
                        sincospi((double)((k * n) & 511) / 256.0, &sn, &cn);
                        re += vs[h] * cn;
                        im -= vs[h] * sn;
                    }
                }
                re = block_sum(re, sRed);
                im = block_sum(im, sRed);
                e += (double)mw[t] * (re * re + im * im);
            }
            if (tid == 0) sStage[rf * stage_pitch + j] = (float)log(fmax(e, (double)kEps));
        }
        if (n_rep > 0 || sRepair[0] != 0) {
            __syncthreads();
            if (tid == 0) sRepair[0] = 0;
        }
        // column sums for the CMN, taken from the tile AFTER the repair so that both CMN paths (this CTA's registers
        // or cmn_kernel's re-read of the stored rows) add the same values in the same order: frames f, f+16, ...
        // per lane, then a tree over the 16 frame lanes
        if (A.mean_nor && f0 + mf < f_end) {
#pragma unroll
            for (int s = 0; s < kMaxSlots; ++s) {
                if (s >= n_slots) break;
                const int j = sSlot[mg * kMaxSlots + s];
                if (j >= 0) colsum[s] += sStage[mf * stage_pitch + j];
            }
        }
        // =============================================================== store the round's tile (contiguous rows)
        {
            const int rows = min(kFramesPerRound, f_end - f0);
            float *dst = orow + (int64_t)f0 * n_mels;
            // thread (row = tid / 16, 16 column lanes): 64-byte runs per row, no index division
            for (int r = tid >> 4; r < rows; r += kThreads / 16)
                for (int col = tid & 15; col < n_mels; col += 16) dst[r * n_mels + col] = sStage[r * stage_pitch + col];
        }
        // the next round's FFT phase only touches sXp/sPw/sEf; sStage is rewritten after its barrier
    }

    if (!A.mean_nor) return;
    // ---- column sums over this CTA's frames: reduce the 16 frame lanes of each group
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
        if (s >= n_slots) break;
        float v = colsum[s];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
        const int j = sSlot[mg * kMaxSlots + s];
        if (mf == 0 && j >= 0) sColMean[j] = v;
    }
    __syncthreads();
    if (!A.fused) return;      // frame-range CTAs: cmn_kernel does the normalisation
    // ---- utterance CMN (processor.py:156-157): second pass over this CTA's own rows (L2-resident)
    for (int j = tid; j < n_mels; j += kThreads) sColMean[j] = sColMean[j] / (float)A.m;
    __syncthreads();
    for (int r = tid >> 4; r < A.m; r += kThreads / 16)
        for (int col = tid & 15; col < n_mels; col += 16) orow[r * n_mels + col] -= sColMean[col];
}

// second pass for utterances split over several CTAs: out[b, :, c] -= mean over frames.  Fixed summation order
// (16 frame lanes per column, tree over lanes), so results do not depend on scheduling.
__global__ void __launch_bounds__(256)
cmn_kernel(float *__restrict__ out, int m, int n_mels) {
    __shared__ float mean[kMaxMels];
    float *orow = out + (int64_t)blockIdx.x * m * n_mels;
    const int lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4;      // 16 groups x 16 frame lanes
    for (int c = grp; c < n_mels; c += 16) {
        float s = 0.f;
        for (int fr = lane16; fr < m; fr += 16) s += orow[(size_t)fr * n_mels + c];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
        if (lane16 == 0) mean[c] = s / (float)m;
    }
    __syncthreads();
    for (int r = grp; r < m; r += 16)
        for (int col = lane16; col < n_mels; col += 16) orow[(size_t)r * n_mels + col] -= mean[col];
}

// ------------------------------------------------------------------ host tables
struct HostState {
    std::mutex mu;
    float repair_theta = 5e-4f;
    int repair_cap = 8;
    std::vector<double> window;           // 400 override (empty = built-in povey window in float64)
    std::vector<float> mel;               // [n_mels, 256] override (empty = built-in)
    int override_mels = 0;
    struct Dev { int device; int n_mels; FbankTables *ptr; uint64_t gen; int n_w; };
    std::vector<Dev> devs;
    uint64_t gen = 1;
};
HostState &state() {
    static HostState s;
    return s;
}

// kaldi.py:98-100 in float64 (the precision of the 'truth' run of the reference)
void builtin_window(double *w) {
    for (int i = 0; i < kFrameLen; ++i) {
        const double h = 0.5 - 0.5 * std::cos(2.0 * M_PI * i / (kFrameLen - 1));
        w[i] = std::pow(h, 0.85);
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
            mel[(size_t)j * kBins + k] = (float)std::fmax(0.0, std::fmin(up, down));
        }
    }
}

int build_tables(int n_mels, FbankTables &t) {
    HostState &s = state();
    std::vector<float> mel;
    if (s.override_mels == n_mels && !s.mel.empty()) mel = s.mel;
    else builtin_mel(n_mels, mel);
    if (!s.window.empty()) std::copy(s.window.begin(), s.window.end(), t.win64);
    else builtin_window(t.win64);
    // the split X = H B + C needs a window that vanishes at both ends (true for povey/hann/hamming^... = 0 there)
    for (int n = 0; n < kFrameLen; ++n) {
        const double wn = t.win64[n], wn1 = (n + 1 < kFrameLen) ? t.win64[n + 1] : 0.0;
        t.T[n] = make_float2((float)wn1, (float)((double)kScaleC * (wn - wn1)));
    }
    for (int k1 = 0; k1 < 32; ++k1)
        for (int c = 0; c < 16; ++c) {
            const double a = -2.0 * M_PI * (double)(c * k1) / 512.0;
            t.tw[k1 * 16 + c] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    std::vector<double> H2(kBins);
    for (int k = 0; k < kBins; ++k) {
        const double a = -2.0 * M_PI * k / 512.0;
        const double hr = 1.0 - kPreemph * std::cos(a), hi = -kPreemph * std::sin(a);
        t.Hh[k] = make_float2((float)(0.5 * hr), (float)(0.5 * hi));
        H2[k] = hr * hr + hi * hi;
    }
    // exact runs per filter
    int nnz = 0;
    for (int j = 0; j < n_mels; ++j) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k)
            if (mel[(size_t)j * kBins + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        const int cnt = first < 0 ? 0 : last - first + 1;
        if (nnz + cnt > 4096) {
            set_error("mel bank too dense (%d non-zeros > 4096)", nnz + cnt);
            return SPK_ERR_UNSUPPORTED;
        }
        t.mel_start[j] = first < 0 ? 0 : first;
        t.mel_count[j] = cnt;
        t.mel_woff[j] = nnz;
        double ge = 0.0;
        for (int k = 0; k < cnt; ++k) {
            t.mel_w[nnz + k] = mel[(size_t)j * kBins + first + k];
            ge += (double)t.mel_w[nnz + k] * H2[first + k];
        }
        t.gE[j] = (float)ge;
        nnz += cnt;
    }
    // filter groups for the mel phase: group g takes filters 32 u + g and 32 u + 31 - g (a snake over the
    // monotonically widening filters keeps the groups balanced); the two groups of a warp share a trip count
    const int n_slots = (n_mels + kGroups - 1) / kGroups;
    int n_w = 0;
    for (int w = 0; w < kGroups / 2; ++w)
        for (int sl = 0; sl < kMaxSlots; ++sl) {
            int trip = 0;
            int fj[2];
            for (int h = 0; h < 2; ++h) {
                const int g = 2 * w + h;
                int j = 32 * (sl / 2) + ((sl & 1) ? 31 - g : g);
                if (sl >= n_slots || j >= n_mels) j = -1;
                fj[h] = j;
                if (j >= 0) trip = std::max(trip, t.mel_count[j]);
            }
            trip = (trip + 3) & ~3;          // the tap loop runs 4 taps per step (one 16-byte weight load)
            t.slot_trip[w * kMaxSlots + sl] = trip;
            for (int h = 0; h < 2; ++h) {
                const int q = (2 * w + h) * kMaxSlots + sl, j = fj[h];
                t.slot_filter[q] = j;
                t.slot_start[q] = j >= 0 ? t.mel_start[j] : 0;
                t.slot_woff[q] = n_w;
                if (n_w + trip > kMaxW) {
                    set_error("mel bank too dense for the grouped layout");
                    return SPK_ERR_UNSUPPORTED;
                }
                for (int k = 0; k < trip; ++k)
                    t.melw[n_w + k] = (j >= 0 && k < t.mel_count[j]) ? t.mel_w[t.mel_woff[j] + k] : 0.f;
                n_w += trip;
            }
        }
    t.n_mels = n_mels;
    t.n_slots = n_slots;
    t.n_w = n_w;
    return SPK_OK;
}

int get_tables(int n_mels, const FbankTables **out, int *n_w) {
    HostState &s = state();
    std::lock_guard<std::mutex> lk(s.mu);
    int dev = 0;
    SPK_CUDA_OK(cudaGetDevice(&dev));
    for (auto &d : s.devs)
        if (d.device == dev && d.n_mels == n_mels && d.gen == s.gen) {
            *out = d.ptr;
            *n_w = d.n_w;
            return SPK_OK;
        }
    static FbankTables host_t;   // guarded by s.mu
    int rc = build_tables(n_mels, host_t);
    if (rc != SPK_OK) return rc;
    FbankTables *dptr = nullptr;
    SPK_CUDA_OK(cudaMalloc(&dptr, sizeof(FbankTables)));
    SPK_CUDA_OK(cudaMemcpy(dptr, &host_t, sizeof(FbankTables), cudaMemcpyHostToDevice));
    s.devs.push_back({dev, n_mels, dptr, s.gen, host_t.n_w});
    *out = dptr;
    *n_w = host_t.n_w;
    return SPK_OK;
}

size_t smem_bytes(int n_mels, int n_w) {
    size_t b = 0;
    b += (416 + 512 + kBins + kFramesPerRound * kXpFloat2) * sizeof(float2);
    b += (size_t)(kPwRows * kPwPitch + kFramesPerRound * (n_mels + 1) + ((n_w + 3) & ~3)) * sizeof(float);
    b += (size_t)(3 * kGroups * kMaxSlots + (kGroups / 2) * kMaxSlots) * sizeof(int);
    b += (size_t)(kFramesPerRound + 2 * kMaxMels) * sizeof(float);
    b += (size_t)(2 + kRepairMax) * sizeof(int);
    b = (b + 7) & ~size_t(7);
    b += 8 * sizeof(double);
    return b;
}

template <typename SampleT>
int launch(const void *wav, const int64_t *starts, const int32_t *lens, const int32_t *phases, int64_t B,
           int64_t n_samples, int64_t stride,
           float *out, int n_mels, int mean_nor, cudaStream_t stream) {
    NvtxRange range("spk_fbank");
    SPK_REQUIRE(B >= 0, "negative batch");
    SPK_REQUIRE(B == 0 || (wav != nullptr && out != nullptr), "null buffer");
    // kaldi.py:142: assert 2 <= window_size <= len(waveform)
    SPK_REQUIRE(n_samples >= kFrameLen, "choose a window size %d that is [2, %lld]", kFrameLen, (long long)n_samples);
    SPK_REQUIRE(starts != nullptr || stride >= n_samples, "wav_stride %lld < n_samples %lld", (long long)stride,
                (long long)n_samples);
    SPK_REQUIRE(n_mels > 3 && n_mels <= kMaxMels, "n_mels %d out of range (3,%d]", n_mels, kMaxMels);
    if (B == 0) return SPK_OK;
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    const FbankTables *tab = nullptr;
    int n_w = 0;
    rc = get_tables(n_mels, &tab, &n_w);
    if (rc != SPK_OK) return rc;
    const int64_t m64 = spk_fbank_num_frames(n_samples);
    SPK_REQUIRE(m64 < (1ll << 30), "too many frames");
    SPK_REQUIRE(B < (1ll << 31), "batch too large");
    const int m = (int)m64;

    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        const int cap = (int)smem_bytes(kMaxMels, kMaxW);
        for (auto fn : {(const void *)fbank_kernel<float, false>, (const void *)fbank_kernel<float, true>,
                        (const void *)fbank_kernel<int16_t, false>, (const void *)fbank_kernel<int16_t, true>})
            if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    });
    SPK_CUDA_OK(attr_err);
    const size_t smem = smem_bytes(n_mels, n_w);

    // one CTA per utterance (fused CMN) when that already fills the GPU twice over; otherwise frame ranges
    const int64_t want_ctas = 4ll * sm_count();
    int64_t fpc = (B * m + want_ctas - 1) / want_ctas;
    fpc = std::max<int64_t>(2 * kFramesPerRound, (fpc + kFramesPerRound - 1) / kFramesPerRound * kFramesPerRound);
    FbankArgs a{};
    a.wav = wav; a.starts = starts; a.lens = lens; a.phases = phases; a.stride = stride; a.n_samples = n_samples; a.out = out;
    a.m = m; a.n_mels = n_mels; a.mean_nor = mean_nor; a.tab = tab;
    a.repair_theta = state().repair_theta; a.repair_cap = state().repair_cap;
    a.fused = fpc >= m;
    a.frames_per_cta = a.fused ? m : (int)fpc;
    const int chunks = a.fused ? 1 : (m + a.frames_per_cta - 1) / a.frames_per_cta;
    SPK_REQUIRE(chunks <= 65535, "utterance too long (%d frames)", m);
    dim3 grid((unsigned)B, (unsigned)chunks);
    if (lens != nullptr) fbank_kernel<SampleT, true><<<grid, kThreads, smem, stream>>>(a);       // windows: circle_pad possible
    else fbank_kernel<SampleT, false><<<grid, kThreads, smem, stream>>>(a);
    rc = check_launch("fbank_kernel");
    if (rc != SPK_OK) return rc;
    if (mean_nor && !a.fused) {
        cmn_kernel<<<(unsigned)B, 256, 0, stream>>>(out, m, n_mels);
        rc = check_launch("cmn_kernel");
    }
    return rc;
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
        SPK_REQUIRE(window400[0] == 0.f && window400[kFrameLen - 1] == 0.f,
                    "the window must vanish at both ends (the povey window does)");
        s.window.assign(window400, window400 + kFrameLen);
    } else {
        s.window.clear();
    }
    s.gen++;   // invalidate device copies (old ones are leaked on purpose: a few KB, rare)
    return SPK_OK;
}

extern "C" int spk_fbank_set_repair(float theta, int cap) {
    SPK_REQUIRE(theta >= 0.f && theta <= 1.f && cap >= 0 && cap <= kRepairMax, "theta in [0,1], cap in [0,%d]", kRepairMax);
    HostState &s = state();
    std::lock_guard<std::mutex> lk(s.mu);
    s.repair_theta = theta;
    s.repair_cap = cap;
    return SPK_OK;
}

extern "C" int spk_fbank_f32(const float *wav, int64_t B, int64_t n_samples, int64_t wav_stride, float *out, int n_mels,
                             int mean_nor, void *stream) {
    return launch<float>(wav, nullptr, nullptr, nullptr, B, n_samples, wav_stride, out, n_mels, mean_nor,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int spk_fbank_i16(const int16_t *wav, int64_t B, int64_t n_samples, int64_t wav_stride, float *out, int n_mels,
                             int mean_nor, void *stream) {
    return launch<int16_t>(wav, nullptr, nullptr, nullptr, B, n_samples, wav_stride, out, n_mels, mean_nor,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int spk_fbank_windows(const void *wav, int is_int16, int64_t n_total, const int64_t *starts, const int32_t *lens,
                                 const int32_t *phases, int64_t B, int64_t n_samples, float *out, int n_mels, int mean_nor,
                                 void *stream) {
    SPK_REQUIRE(B == 0 || (starts != nullptr && lens != nullptr), "null window table");
    SPK_REQUIRE(n_total > 0, "empty recording");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_int16) return launch<int16_t>(wav, starts, lens, phases, B, n_samples, 0, out, n_mels, mean_nor, s);
    return launch<float>(wav, starts, lens, phases, B, n_samples, 0, out, n_mels, mean_nor, s);
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
