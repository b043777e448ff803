/* b200spk.h - C ABI of libb200spk.so, the sm_100a implementation of 3D-Speaker's
 * embedding-extraction hot path (waveform -> Kaldi fbank+CMN -> speaker network forward ->
 * cosine affinity + spectral clustering).
 *
 * The reference has no FFI for this path: the boundary is a set of Python objects resolved by
 * dotted name (speakerlab/utils/builder.py:9-12,52-60).  Each entry point below names the
 * reference interface whose arithmetic it replaces; the Python mirrors of those interfaces
 * (3d-speaker_b200/b200spk/) call these through ctypes with tensor.data_ptr() + the current
 * CUDA stream.  See INTEGRATION.md for the reference-side binding.
 *
 * Conventions: plain pointers and sizes only; every function returns SPK_OK (0) or a negative
 * SPK_ERR_* code and never throws; spk_last_error() gives the message for the calling thread.
 * Device entry points are asynchronous on the caller's stream (a cudaStream_t passed as void*)
 * and never synchronise; the caller owns every device buffer including the workspace.
 * *_host entry points take HOST buffers, do their own H2D/D2H and return when results are in
 * host memory.  There is no CPU fallback anywhere: without an sm_100 device the compute entry
 * points return SPK_ERR_NO_DEVICE.
 */
#ifndef B200SPK_H_
#define B200SPK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPK_OK                 0
#define SPK_ERR_INVALID       -1   /* bad argument (the reference would trip a Python assert) */
#define SPK_ERR_CUDA          -2   /* CUDA runtime / launch failure                          */
#define SPK_ERR_UNSUPPORTED   -3   /* shape or mode not implemented                          */
#define SPK_ERR_WORKSPACE     -4   /* caller-provided workspace too small                    */
#define SPK_ERR_NO_DEVICE     -5   /* no sm_100 CUDA device                                  */
#define SPK_ERR_KERNEL        -6   /* a kernel reported an internal error (watchdog)         */

#define SPK_ABI_VERSION        1

/* ------------------------------------------------------------------ library */
int         spk_abi_version(void);
const char *spk_last_error(void);
/* SPK_OK when device 'dev' exists and is compute capability 10.x */
int         spk_device_check(int dev);

/* ------------------------------------------------------------------ front end
 * Replaces FBank.__call__ (speakerlab/process/processor.py:143-158) =
 * torchaudio.compliance.kaldi.fbank(num_mel_bins, sample_frequency=16000, dither=0)
 * (+ utterance CMN when mean_nor), in its batched torch.vmap form
 * (speakerlab/bin/infer_diarization.py:634).
 */
/* 1 + (n-400)/160, or 0 when n < 400 (kaldi.py:67) */
int64_t spk_fbank_num_frames(int64_t n_samples);
/* Optional: replace the built-in mel bank [n_mels,256] (the Python mirror passes the bank built with
 * torchaudio's own float32 arithmetic, kaldi.py:436-511, so the weights agree to the last bit) and the
 * window [400] (must be 0 at both ends; NULL = the povey window evaluated in float64, which is what the
 * float64 run of the reference uses).  Host pointers; copied.  NULL keeps the built-in table. */
int spk_fbank_set_tables(const float *window400, const float *mel_bank, int n_mels);
/* Accuracy knob.  Mel cells whose energy is below theta x (the energy white noise with the frame's power would
 * put there) are where float32 cannot deliver 1e-4 in the log (the reference's own float32 path is off by 1e-3
 * there); up to cap such cells per 16-frame round are recomputed in float64.  Defaults theta = 5e-4, cap = 8
 * (a few cells per million on noise-like audio, none on speech); cap = 0 disables the repair, cap <= 256. */
int spk_fbank_set_repair(float theta, int cap);
/* wav: device, B rows of n_samples floats in [-1,1] scale, row stride wav_stride (elements).
 * out: device, [B, m, n_mels] contiguous.  n_samples >= 400 (kaldi.py:142). */
int spk_fbank_f32(const float *wav, int64_t B, int64_t n_samples, int64_t wav_stride,
                  float *out, int n_mels, int mean_nor, void *stream);
/* Same on int16 PCM rows: samples are scaled by 1/32768 on load (speakerlab/utils/fileio.py:115-117,
 * load_audio's int16 -> float conversion), so recordings can cross PCIe at 2 bytes per sample. */
int spk_fbank_i16(const int16_t *wav, int64_t B, int64_t n_samples, int64_t wav_stride,
                  float *out, int n_mels, int mean_nor, void *stream);
/* Window mode: the batch rows are pieces of recordings resident in ONE device buffer (float32 or int16, n_total
 * samples).  Sample i of row b is wav[starts[b] + (phases[b] + i) % lens[b]]:
 *  - diarization sub-segments (speakerlab/bin/infer_diarization.py:621-627): starts = window start, lens = window
 *    length, phases = NULL -> a window shorter than n_samples is circle-padded (speakerlab/utils/utils.py:232-238);
 *  - bulk extraction chunks (speakerlab/bin/infer_sv_batch.py:388-412): starts = start of the (truncated) recording,
 *    lens = its length, phases = chunk index * n_samples -> the recording is circle-padded to a whole number of
 *    chunks and sliced, without ever materialising [B, n_samples].
 * starts (int64), lens and phases (int32) are device arrays; [starts[b], starts[b] + lens[b]) must lie inside the
 * buffer. */
int spk_fbank_windows(const void *wav, int is_int16, int64_t n_total, const int64_t *starts, const int32_t *lens,
                      const int32_t *phases, int64_t B, int64_t n_samples, float *out, int n_mels, int mean_nor,
                      void *stream);
/* Same as spk_fbank_f32, host buffers in / host buffers out (pageable or pinned). */
int spk_fbank_host_f32(const float *wav, int64_t B, int64_t n_samples, int64_t wav_stride,
                       float *out, int n_mels, int mean_nor);

/* ------------------------------------------------------------------ network forward
 * Replaces nn.Module.forward of the embedding models
 * (CAMPPlus.forward speakerlab/models/campplus/DTDNN.py:111-115;
 *  ERes2NetV2.forward speakerlab/models/eres2net/ERes2NetV2.py:235-254) in eval mode.
 * The host mirror folds eval-mode BatchNorm into per-channel scale/shift, repacks conv
 * weights to [Cout][KH][KW][Cin] and describes the network as a list of fused ops over
 * channels-last activation buffers; the library owns the device copy of the parameters and
 * runs the op list with its own kernels.
 */
typedef struct spk_model spk_model_t;

/* SPK_PREC_BF16: bf16 activations and weights, fp32 accumulation (tcgen05 kind::f16).
 * SPK_PREC_F32:  fp32 activations and weights; convs run as 3xTF32 split products on the tensor cores
 *                with fp32 register sums (csrc/conv_f32x3.cu): the result is fp32-accurate (about 2e-7 per
 *                conv against float64), the mode whose embeddings match the reference's CPU forward to
 *                1e-4 rel-L2.  SPK_NO_F32X3=1 in the environment moves them to CUDA cores. */
enum { SPK_PREC_F32 = 0, SPK_PREC_BF16 = 1 };
enum { SPK_DT_F32 = 0, SPK_DT_BF16 = 1 };

/* activation buffer: elems per segment; ids 0 (input feats [T,F] f32) and 1 (embedding f32)
 * are bound to the caller's pointers at forward time, the rest live in the workspace */
typedef struct {
    int64_t elems;
    int32_t dtype;     /* SPK_DT_* */
    int32_t reserved;
} spk_buf_t;

enum {
    SPK_OP_STEM       = 1,  /* conv2d 1->Cout 3x3 pad 1 on feats[T,F] read as image [F,T] + affine + relu */
    SPK_OP_CONV       = 2,  /* generic NHWC conv as implicit GEMM with fused prologue/epilogue        */
    SPK_OP_CAM_GATE   = 3,  /* CAM context: mean_T + seg-mean -> 1x1 -> relu -> 1x1 -> sigmoid        */
    SPK_OP_STATS_POOL = 4,  /* mean / std over positions                                             */
    SPK_OP_AFF_BLEND  = 5,  /* x*g + y*(2-g), g = 1+tanh(z)  (fusion.py:22-28); z = none: x + y       */
    SPK_OP_CAM_LOCAL  = 6,  /* whole CAMLayer: dilated k=3 conv x context gate (layers.py:93-99); CONV fields +
                               aux = w1,b1,w2,b2, iaux = hidden, seg_len, id(w1^T), id(w2^T); gate_buf = scratch for the unfused path */
    SPK_OP_SE_SCALE   = 7,  /* out = in * gate[b, c] + res   (SEBlock + block residual, ECAPA_TDNN.py:222,345)    */
    SPK_OP_STEM_BLOCK = 9,  /* bf16 only: STEM fused with the first BasicResBlock's conv1 (3x3 stride (2,1)) and 1x1 shortcut
                               (DTDNN.py:39-48, layers.py:221-253), the stem output is never stored.  in_buf = feats, w /
                               epi_scale / epi_shift = stem conv + BN; out_buf = conv1 output with aux[0] = conv1 weights,
                               aux[1], aux[2] = its BN scale / shift (ReLU); res_buf (res_ld, res_choff) = shortcut OUTPUT with
                               aux[3] = shortcut weights, iaux[0], iaux[1] = its BN scale / shift; H = F, W = T, Ho = F / 2 */
    SPK_OP_ASP_POOL   = 8   /* softmax over positions of in (logits), weighted mean/std of res (ECAPA_TDNN.py:279-285);
                               faux[0] = variance floor */
};
/* SPK_OP_CONV extras for ECAPA-TDNN: aux[0], aux[1] = per-channel scale/shift applied after the activation
 * (TDNNBlock is conv -> ReLU -> BN), iaux[0] = activation after that affine, iaux[1] = 1: reflect padding along W,
 * iaux[2] = 1: gate_buf is a per-segment additive bias applied before the activation.
 * SPK_OP_CAM_GATE: iaux[2] = 1: squeeze-excitation mode (context = mean over all positions only).
 * SPK_OP_STATS_POOL: faux[1] > 0: std = sqrt(max(var, faux[1])) (ASP global context).
 * SPK_OP_AFF_BLEND: res_buf = -1: plain (dtype-converting) copy. */
enum { SPK_ACT_NONE = 0, SPK_ACT_RELU = 1, SPK_ACT_CLAMP20 = 2, SPK_ACT_SILU = 3, SPK_ACT_TANH = 4 };

typedef struct {
    int32_t kind;
    int32_t in_buf, in_ld, in_choff;       /* buffer id, elems per pixel, first channel   */
    int32_t out_buf, out_ld, out_choff;
    int32_t res_buf, res_ld, res_choff;    /* residual added before the activation, or -1 */
    int32_t gate_buf, gate_win;            /* per-(segment, window, channel) multiplier, or -1 */
    int32_t H, W, Cin, Ho, Wo, Cout;
    int32_t KH, KW, sh, sw, ph, pw, dh, dw;
    int32_t w;                             /* parameter ids (spk_model_add_param), -1 = none */
    int32_t pro_scale, pro_shift, pro_relu;/* per-input-channel affine (+relu) on the A operand */
    int32_t epi_scale, epi_shift, act;     /* per-output-channel affine, activation          */
    int32_t aux[4];                        /* CAM_GATE: w1,b1,w2,b2                          */
    int32_t iaux[4];                       /* CAM_GATE: hidden, seg_len; STATS_POOL: unbiased */
    float   faux[2];                       /* STATS_POOL: eps inside sqrt                    */
    int32_t phase;                         /* 0: runs per fine sub-batch; 1: per coarse sub-batch */
    int32_t reserved;                      /* SPK_OP_CONV, stride-1 1x1 (TMA GEMM): L2 hints, bit 0 = the input is a stream (load
                                              evict_first), bit 1 = keep the output in L2 for the next op (store evict_last) */
} spk_op_t;

int     spk_model_create(spk_model_t **out, int precision);
int     spk_model_destroy(spk_model_t *m);
/* copy n floats (host) into the model's device arena; returns the parameter id (>=0) */
int64_t spk_model_add_param(spk_model_t *m, const float *host, int64_t n);
/* register the op list for segments of T frames (replaces an earlier program for the same T) */
int     spk_model_set_program(spk_model_t *m, int64_t T, const spk_buf_t *bufs, int32_t n_bufs,
                              const spk_op_t *ops, int32_t n_ops);
/* bytes of workspace spk_model_forward needs for the given sub-batch sizes */
int64_t spk_model_workspace_bytes(spk_model_t *m, int64_t T, int64_t chunk, int64_t fine_chunk);
/* feats: device [B,T,F] f32 contiguous; emb: device [B,E] f32.  B is walked in sub-batches of
 * at most 'chunk' segments; inside each, the phase-0 ops (large per-segment activations, e.g.
 * the 2-D front of CAM++) run over sub-batches of 'fine_chunk' segments so their
 * producer->consumer traffic stays inside L2, then the phase-1 ops run once over the whole
 * sub-batch so their grids fill the GPU.  fine_chunk <= 0 means fine_chunk = chunk. */
int     spk_model_forward(spk_model_t *m, int64_t T, const float *feats, int64_t B, float *emb,
                          void *workspace, int64_t workspace_bytes, int64_t chunk, int64_t fine_chunk,
                          void *stream);
/* debugging / per-layer parity: copy the first n elements of workspace buffer 'buf_id' of the
 * LAST sub-batch into dst (device pointer, f32; bf16 buffers are widened) */
int     spk_model_read_buffer(spk_model_t *m, int64_t T, int32_t buf_id, int64_t chunk, int64_t fine_chunk,
                              const void *workspace, float *dst, int64_t n, void *stream);
/* kernels launched by spk_* calls since process start (bench.py's gpu_launches) */
int64_t spk_launch_count(void);
/* a host layer that replays captured launches (CUDA graph) adds them here so the counter stays truthful */
void    spk_add_launches(int64_t n);

/* ------------------------------------------------------------------ back end
 * Replaces SpectralCluster.__call__ (speakerlab/process/cluster.py:35-57).
 */
/* X: device [N,D] f32 (rows need not be normalised).  Writes the unnormalised Laplacian
 * L = D - M of the p-pruned, symmetrised cosine affinity (cluster.py:59-84) as dense f32 with row
 * pitch Np = N rounded up to 16 (L must hold Np*Np floats; rows/cols >= N are not written).
 * keep = N - n_elems entries survive per row (cluster.py:67-68); ties at the cut keep the higher
 * column index (what a stable argsort zeroes last). */
int spk_affinity_laplacian(const float *X, int64_t N, int64_t D, int64_t keep, float *L,
                           void *workspace, int64_t workspace_bytes, void *stream);
int64_t spk_affinity_workspace_bytes(int64_t N, int64_t D);
/* k <= 32 smallest eigenpairs of the symmetric PSD L written by spk_affinity_laplacian (row pitch
 * Np) by Lanczos with full re-orthogonalisation on sigma*I - L (replaces scipy eigsh(which='SM'),
 * cluster.py:90).  evals: host [k] ascending; evecs: device [N,k] row-major, column j =
 * eigenvector j.  Synchronises the stream.  Returns the Krylov dimension used (> 0) or an error. */
int spk_eig_smallest(const float *L, int64_t N, int32_t k, float *evals_host, float *evecs,
                     void *workspace, int64_t workspace_bytes, void *stream);
int64_t spk_eig_workspace_bytes(int64_t N, int32_t k);
/* The same eigensolver in two steps, for hosts that bring their own tridiagonal solver (the
 * Python mirror uses LAPACK through scipy.linalg.eigh_tridiagonal): spk_lanczos_extend grows the
 * Krylov basis kept in the workspace from m_from to m_to vectors (m_from = 0 starts a run) and
 * returns the tridiagonal of sigma*I - L (alpha/beta, host) and the shift sigma;
 * spk_lanczos_ritz turns S (host [m,k], eigenvectors of that tridiagonal) into evecs [N,k]. */
int spk_lanczos_extend(const float *L, int64_t N, int32_t k, int32_t m_from, int32_t m_to, float *alpha_host,
                       float *beta_host, float *sigma_host, void *workspace, int64_t workspace_bytes,
                       void *stream);
int spk_lanczos_ritz(int64_t N, int32_t k, int32_t m, const float *S_host, float *evecs, void *workspace,
                     int64_t workspace_bytes, void *stream);
int32_t spk_lanczos_max_dim(int64_t N, int32_t k);
/* Average-linkage agglomerative clustering on the distance -cos, cut where the linkage distance exceeds -cos_thr
 * (AHCluster.__call__, speakerlab/process/cluster.py:139-156; the reference's default back end for short
 * recordings, cluster.py:178-181, 189-190).  X: device [N,D] f32; labels: device int32 [N], clusters numbered
 * by their smallest member index.  Returns the number of clusters (>= 1) or an error (synchronises the stream). */
int64_t spk_ahc_workspace_bytes(int64_t N, int64_t D);
int spk_ahc(const float *X, int64_t N, int64_t D, float cos_thr, int32_t *labels, void *workspace,
            int64_t workspace_bytes, void *stream);
/* Lloyd k-means on device points [N,d] f32 from caller-provided initial centres (host [k,d]);
 * stops when no label changes or the summed squared centre shift is <= tol.  labels: device
 * int32 [N].  Deterministic (fixed-order reductions).  Returns iterations run or an error. */
int spk_kmeans(const float *pts, int64_t N, int32_t d, int32_t k, const float *init_centres_host,
               int32_t max_iter, float tol, int32_t *labels, float *inertia_host,
               void *workspace, int64_t workspace_bytes, void *stream);
int64_t spk_kmeans_workspace_bytes(int64_t N, int32_t d, int32_t k);
/* per-recording embedding = mean of its chunk embeddings (speakerlab/bin/infer_sv_batch.py:313-315):
 * out[w] = mean of E[pos[w] .. pos[w+1]) ; E device [N,D] f32, pos device int32 [n_wav + 1] ascending, out [n_wav,D] */
int spk_segment_mean(const float *E, int64_t D, const int32_t *pos, int64_t n_wav, float *out, void *stream);
/* cosine score of trial pairs: out[i] = cos(E[a[i]], E[b[i]])
 * (speakerlab/bin/compute_score_metrics.py:113-114) */
int spk_cosine_pairs(const float *E, int64_t N, int64_t D, const int32_t *a, const int32_t *b,
                     int64_t n_pairs, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SPK_H_ */
