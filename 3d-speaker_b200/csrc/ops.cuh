// Internal launch interface between the op-list interpreter (model.cu) and the kernels.
#pragma once
#include "common.cuh"

namespace spk {

struct ConvArgs {
    const void *x;        // [B,H,W,in_ld] channels-last, first channel in_choff
    const void *w;        // packed [Cout][KH][KW][Cin] (f32 for SIMT, bf16 for the tcgen05 path)
    void *y;              // [B,Ho,Wo,out_ld]
    const void *res;      // optional residual [B,Ho,Wo,res_ld]
    const float *gate;    // optional [B, gate_nwin, Cout]
    const float *pro_scale, *pro_shift, *epi_scale, *epi_shift;
    int B, H, W, Cin, Ho, Wo, Cout;
    int KH, KW, sh, sw, ph, pw, dh, dw;
    int in_ld, in_choff, out_ld, out_choff, res_ld, res_choff;
    int gate_win, gate_nwin, pro_relu, act;
    int K;                // KH*KW*Cin
    long long M;          // B*Ho*Wo
    // ECAPA-TDNN extras (speakerlab/models/ecapa_tdnn/ECAPA_TDNN.py): conv -> act -> BN blocks, reflect padding, ASP
    const float *post_scale, *post_shift;   // per-output-channel affine applied AFTER the activation (TDNNBlock: conv-ReLU-BN)
    int post_act;         // activation after that affine (tanh in the ASP attention)
    int pad_reflect;      // out-of-range columns mirror (F.pad mode='reflect') instead of reading zero; W axis only
    int gate_additive;    // gate[b, window, n] is ADDED before the activation (per-segment bias) instead of multiplied after
    const void *pro_scale_bf, *pro_shift_bf;   // bf16 copies of pro_scale/pro_shift owned by the model (tensor-core prologue)
    int l2_flags;         // conv_gemm: bit 0 = load the input evict_first (a stream), bit 1 = store the output evict_last (keep in L2)
};

// fp32-accumulate CUDA-core implicit GEMM (exact-fp32 mode and odd shapes)
int launch_conv_simt(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype, cudaStream_t s);
// generic tcgen05 / TMEM implicit GEMM, bf16 operands (conv_tc2.cu: TMA weights, double-buffered gather of A)
bool conv_tc_supported(const ConvArgs &a, int in_dtype);
int launch_conv_tc(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s);
int launch_conv_tc2(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s);
// TMA-fed GEMM for stride-1 1x1 convs with the BN-ReLU prologue applied in shared memory (conv_gemm.cu)
bool conv_gemm_supported(const ConvArgs &a, int in_dtype);
int launch_conv_gemm(const ConvArgs &a, int out_dtype, int res_dtype, cudaStream_t s);
// fp32 precision mode on the tensor cores (conv_f32x3.cu): 3xTF32 split products, fp32 in / out / residual.
// a.w = [2][Cout][K] fp32, the TF32 heads then tails of the packed weights (launch_split_tf32)
bool conv_f32x3_supported(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype);
int launch_conv_f32x3(const ConvArgs &a, cudaStream_t s);
int launch_split_tf32(const float *src, float *hi, float *lo, long long n, cudaStream_t s);
// slab kernel for the Cin=Cout=32 2-D convs of the CAM++ head (conv_slab.cu)
bool conv_slab_supported(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype);
int launch_conv_slab(const ConvArgs &a, cudaStream_t s);
// third generation of the same (conv_slab3.cu): TMA-staged 64B-swizzled slabs, lean MMA issue
bool conv_slab3_supported(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype);
int launch_conv_slab3(const ConvArgs &a, cudaStream_t s);
// 3x3 stride-1 convs with up to 64 channels, rows of any width (conv_slab4.cu): column parts, 64/128-byte pixels
bool conv_slab4_supported(const ConvArgs &a, int in_dtype, int out_dtype, int res_dtype);
int launch_conv_slab4(const ConvArgs &a, cudaStream_t s);

// fused CAM layer (cam_local.cu): dilated k=3 conv (128->32) + context gate + gating multiply
bool cam_local_supported(const ConvArgs &a, int in_dtype, int out_dtype, int hidden, int seg_len);
// w1t [C][hidden] and w2t [hidden][Cout] are the TRANSPOSED gate weights
int launch_cam_local(const ConvArgs &a, const float *w1t, const float *b1, const float *w2t, const float *b2, int hidden,
                     int seg_len, cudaStream_t s);

struct StemArgs {
    const float *feats;   // [B,T,F]
    const float *w;       // [Cout][3][3]  (kh over F, kw over T)
    const float *scale, *shift;
    void *y;              // [B,F,T,out_ld]
    int B, T, F, Cout, out_ld, out_choff, act;
};
int launch_stem(const StemArgs &a, int out_dtype, cudaStream_t s);

// stem conv + first FCM block entry in one kernel (conv_stem.cu): y1 = relu(bn1(conv3x3 s(2,1)(stem))), y2 = bn_s(conv1x1 s(2,1)(stem)),
// stem = relu(bn0(conv3x3(feats))) never stored
struct StemBlockArgs {
    const float *feats;                 // [B,T,F]
    const float *w0, *s0, *b0;          // stem [32][3][3] (kh over F, kw over T), folded BN
    const __nv_bfloat16 *w1;            // conv1 [32][3][3][32]
    const float *s1, *b1;
    const __nv_bfloat16 *ws;            // shortcut [32][32]
    const float *ss, *bs;
    void *y1, *y2;                      // [B,F/2,T,ld] bf16
    int B, T, F, Ho, y1_ld, y1_choff, y2_ld, y2_choff;
};
bool stem_block_supported(const StemBlockArgs &a);
int launch_stem_block(const StemBlockArgs &a, cudaStream_t s);

struct CamGateArgs {
    const void *x;        // [B,T,in_ld] (post BN-ReLU bottleneck)
    float *gate;          // [B,nwin,Cout]
    const float *w1, *b1, *w2, *b2;   // [hidden,C],[hidden],[Cout,hidden],[Cout]
    int B, T, C, in_ld, in_choff, hidden, Cout, seg_len, nwin;
    int se_mode;          // squeeze-excitation: context = mean over all positions only (ECAPA SEBlock)
};
int launch_cam_gate(const CamGateArgs &a, int in_dtype, cudaStream_t s);

struct StatsPoolArgs {
    const void *x;        // [B,G,P,in_ld]
    float *y;             // [B, 2, G, C]  (mean block then std block)
    int B, G, P, C, in_ld, in_choff, unbiased;
    float eps;            // std = sqrt(var + eps)
    float var_floor;      // > 0: std = sqrt(max(var, var_floor)) instead (ECAPA ASP global context)
};
int launch_stats_pool(const StatsPoolArgs &a, int in_dtype, cudaStream_t s);

struct AffBlendArgs {
    const void *x, *y, *z;   // [M, ld] each; out = x*g + y*(2-g), g = 1 + tanh(z)
    void *out;
    long long M;
    int C, x_ld, x_choff, y_ld, y_choff, z_ld, z_choff, out_ld, out_choff;
};
int launch_aff_blend(const AffBlendArgs &a, int dtype, int out_dtype, cudaStream_t s);

// out = x * gate[b, c] + res  (SEBlock scaling + block residual; res may be null)
struct SeScaleArgs {
    const void *x, *res;
    const float *gate;    // [B, C]
    void *out;
    long long B;
    int P, C, x_ld, x_choff, res_ld, res_choff, out_ld, out_choff;
};
int launch_se_scale(const SeScaleArgs &a, int dtype, int res_dtype, int out_dtype, cudaStream_t s);

// attentive statistics: p = softmax over the P positions of logits[b, :, c]; out[b] = [sum p x | sqrt(max(sum p (x - mean)^2, floor))]
struct AspPoolArgs {
    const void *logits, *x;   // [B, P, ld]
    float *out;               // [B, 2C]
    long long B;
    int P, C, l_ld, l_choff, x_ld, x_choff;
    float var_floor;
};
int launch_asp_pool(const AspPoolArgs &a, int l_dtype, int x_dtype, cudaStream_t s);

// small_ops.cu: cluster split-K linear layer, split-T squeeze-excitation gate, sliced / one-pass pooling for long axes
bool linear_supported(const ConvArgs &a, int in_dtype, int out_dtype);
int launch_linear(const ConvArgs &a, int in_dtype, int out_dtype, cudaStream_t s);
bool reflect_edge_fix_supported(const ConvArgs &a);           // 1-D 'same' conv with reflect padding: zero-padded conv + edge fix
int launch_reflect_edge_fix(const ConvArgs &a, int in_dtype, int out_dtype, cudaStream_t s);     // a.w: [Cout][KW][Cin] in the activation dtype
bool se_gate_cluster_supported(const CamGateArgs &a, int in_dtype);
int launch_se_gate_cluster(const CamGateArgs &a, int in_dtype, cudaStream_t s);
bool stats_pool_sliced_supported(const StatsPoolArgs &a);
int launch_stats_pool_sliced(const StatsPoolArgs &a, int in_dtype, cudaStream_t s);
bool asp_pool_online_supported(const AspPoolArgs &a);
int launch_asp_pool_online(const AspPoolArgs &a, int l_dtype, int x_dtype, cudaStream_t s);

int launch_f32_to_bf16(const float *src, __nv_bfloat16 *dst, long long n, cudaStream_t s);
int launch_widen(const void *src, int dtype, float *dst, long long n, cudaStream_t s);

}  // namespace spk
