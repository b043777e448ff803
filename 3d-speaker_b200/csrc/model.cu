// Network-forward executor: owns the device copy of the parameters and runs the fused-op list
// the host mirror built for a given frame count.  Replaces nn.Module.forward of the embedding
// models (CAMPPlus.forward speakerlab/models/campplus/DTDNN.py:111-115; ERes2NetV2.forward
// speakerlab/models/eres2net/ERes2NetV2.py:235-254) in eval mode.
//
// Memory plan: activations are channels-last tensors carved out of ONE caller-provided
// workspace, sized for a sub-batch ('chunk') of segments so that producer->consumer traffic
// between consecutive ops stays inside the 126 MB L2; the batch is walked chunk by chunk on the
// caller's stream.  Nothing here allocates or synchronises at forward time.
#include <cstdlib>
#include <map>
#include <vector>

#include "ops.cuh"

using namespace spk;

struct spk_program {
    int64_t T = 0;
    std::vector<spk_buf_t> bufs;
    std::vector<spk_op_t> ops;
};

struct spk_model {
    int precision = SPK_PREC_F32;
    int device = 0;
    std::vector<float *> params;
    std::vector<int64_t> param_n;
    std::vector<__nv_bfloat16 *> params_bf16;   // lazily created for tcgen05 convs
    std::vector<float *> params_x3;             // fp32 mode: [2][n] TF32 heads / tails of the conv weights (conv_f32x3.cu)
    std::map<int64_t, spk_program> programs;
};

namespace {

size_t dtype_size(int dt) { return dt == SPK_DT_BF16 ? 2 : 4; }

// A buffer is coarse-scoped when any phase-1 op touches it, else fine-scoped.
void scopes(const spk_program &p, std::vector<char> &coarse) {
    coarse.assign(p.bufs.size(), 0);
    for (const spk_op_t &o : p.ops) {
        if (o.phase == 0) continue;
        for (int id : {o.in_buf, o.out_buf, o.res_buf, o.gate_buf})
            if (id >= 0) coarse[id] = 1;
    }
}

// byte offsets of workspace buffers (ids >= 2) for sub-batches of n (coarse) / f (fine) segments
void plan(const spk_program &p, int64_t n, int64_t f, std::vector<int64_t> &off, std::vector<char> &coarse,
          int64_t &total) {
    scopes(p, coarse);
    off.assign(p.bufs.size(), -1);
    int64_t cur = 0;
    for (size_t i = 2; i < p.bufs.size(); ++i) {
        off[i] = cur;
        cur += align_up(p.bufs[i].elems * (coarse[i] ? n : f) * (int64_t)dtype_size(p.bufs[i].dtype), 1024);
    }
    total = cur;
}

const float *param(const spk_model *m, int id) {
    return (id >= 0 && id < (int)m->params.size()) ? m->params[id] : nullptr;
}

int param_bf16(spk_model *m, int id, const __nv_bfloat16 **out, cudaStream_t s) {
    if (id < 0 || id >= (int)m->params.size()) {
        set_error("bad parameter id %d", id);
        return SPK_ERR_INVALID;
    }
    if (m->params_bf16[id] == nullptr) {
        __nv_bfloat16 *d = nullptr;
        SPK_CUDA_OK(cudaMalloc(&d, (size_t)m->param_n[id] * sizeof(__nv_bfloat16)));
        int rc = launch_f32_to_bf16(m->params[id], d, m->param_n[id], s);
        if (rc != SPK_OK) return rc;
        m->params_bf16[id] = d;
    }
    *out = m->params_bf16[id];
    return SPK_OK;
}

int param_x3(spk_model *m, int id, const float **out, cudaStream_t s) {
    if (id < 0 || id >= (int)m->params.size()) {
        set_error("bad parameter id %d", id);
        return SPK_ERR_INVALID;
    }
    if (m->params_x3[id] == nullptr) {
        float *d = nullptr;
        SPK_CUDA_OK(cudaMalloc(&d, (size_t)m->param_n[id] * 2 * sizeof(float)));
        int rc = launch_split_tf32(m->params[id], d, d + m->param_n[id], m->param_n[id], s);
        if (rc != SPK_OK) return rc;
        m->params_x3[id] = d;
    }
    *out = m->params_x3[id];
    return SPK_OK;
}

// SPK_TRACE_DISPATCH=1: one stderr line per conv op saying which kernel family took it (docs / debugging)
void trace_dispatch(const char *which, const ConvArgs &a) {
    static const bool on = [] { const char *e = getenv("SPK_TRACE_DISPATCH"); return e && e[0] == '1'; }();
    if (on)
        fprintf(stderr, "[spk dispatch] %-10s %dx%dx%d k%dx%d s%d,%d d%d,%d -> %dx%dx%d pro=%d res=%d reflect=%d\n", which, a.H, a.W, a.Cin, a.KH,
                a.KW, a.sh, a.sw, a.dh, a.dw, a.Ho, a.Wo, a.Cout, a.pro_scale != nullptr, a.res != nullptr, a.pad_reflect);
}

int validate(const spk_model *m, const spk_program &p) {
    const int nb = (int)p.bufs.size(), np = (int)m->params.size();
    auto buf_ok = [&](int b, bool allow_none) { return (allow_none && b < 0) || (b >= 0 && b < nb); };
    auto par_ok = [&](int q, bool allow_none) { return (allow_none && q < 0) || (q >= 0 && q < np); };
    for (size_t i = 0; i < p.ops.size(); ++i) {
        const spk_op_t &o = p.ops[i];
        bool ok = buf_ok(o.in_buf, false) && buf_ok(o.out_buf, false) && buf_ok(o.res_buf, true) &&
                  buf_ok(o.gate_buf, true);
        switch (o.kind) {
            case SPK_OP_STEM:
                ok = ok && par_ok(o.w, false) && par_ok(o.epi_scale, true) && par_ok(o.epi_shift, true);
                break;
            case SPK_OP_CONV:
                ok = ok && par_ok(o.w, false) && par_ok(o.pro_scale, true) && par_ok(o.pro_shift, true) &&
                     par_ok(o.epi_scale, true) && par_ok(o.epi_shift, true) &&
                     ((o.pro_scale < 0) == (o.pro_shift < 0)) && ((o.epi_scale < 0) == (o.epi_shift < 0));
                ok = ok && o.KH > 0 && o.KW > 0 && o.sh > 0 && o.sw > 0 && o.Cin > 0 && o.Cout > 0;
                if (ok && m->param_n[o.w] != (int64_t)o.Cout * o.KH * o.KW * o.Cin) ok = false;
                if (ok && o.gate_buf >= 0 && o.gate_win <= 0) ok = false;
                ok = ok && par_ok(o.aux[0], true) && par_ok(o.aux[1], true) && ((o.aux[0] < 0) == (o.aux[1] < 0));
                break;
            case SPK_OP_CAM_GATE:
                for (int j = 0; j < 4; ++j) ok = ok && par_ok(o.aux[j], false);
                ok = ok && o.iaux[0] > 0 && o.iaux[1] > 0;
                break;
            case SPK_OP_CAM_LOCAL:
                for (int j = 0; j < 4; ++j) ok = ok && par_ok(o.aux[j], false);
                ok = ok && par_ok(o.w, false) && o.iaux[0] > 0 && o.iaux[1] > 0 && o.gate_buf >= 0 && o.KH == 1 && o.H == 1 &&
                     par_ok(o.iaux[2], false) && par_ok(o.iaux[3], false);
                if (ok && m->param_n[o.w] != (int64_t)o.Cout * o.KW * o.Cin) ok = false;
                break;
            case SPK_OP_STATS_POOL:
            case SPK_OP_AFF_BLEND:
                break;
            case SPK_OP_STEM_BLOCK:
                ok = ok && m->precision == SPK_PREC_BF16 && par_ok(o.w, false) && par_ok(o.epi_scale, false) && par_ok(o.epi_shift, false) &&
                     par_ok(o.aux[0], false) && par_ok(o.aux[1], false) && par_ok(o.aux[2], false) && par_ok(o.aux[3], false) &&
                     par_ok(o.iaux[0], false) && par_ok(o.iaux[1], false) && o.res_buf >= 0 && o.Cout == 32;
                if (ok && (m->param_n[o.w] != 32 * 9 || m->param_n[o.aux[0]] != 32 * 9 * 32 || m->param_n[o.aux[3]] != 32 * 32)) ok = false;
                break;
            case SPK_OP_SE_SCALE:
                ok = ok && o.gate_buf >= 0;
                break;
            case SPK_OP_ASP_POOL:
                ok = ok && o.res_buf >= 0;
                break;
            default:
                ok = false;
        }
        if (!ok) {
            set_error("op %zu (kind %d) is malformed", i, o.kind);
            return SPK_ERR_INVALID;
        }
    }
    return SPK_OK;
}

}  // namespace

extern "C" int spk_model_create(spk_model_t **out, int precision) {
    SPK_REQUIRE(out != nullptr, "null out");
    SPK_REQUIRE(precision == SPK_PREC_F32 || precision == SPK_PREC_BF16, "bad precision %d", precision);
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    spk_model *m = new spk_model();
    m->precision = precision;
    cudaGetDevice(&m->device);
    *out = m;
    return SPK_OK;
}

extern "C" int spk_model_destroy(spk_model_t *m) {
    if (m == nullptr) return SPK_OK;
    for (float *p : m->params) cudaFree(p);
    for (__nv_bfloat16 *p : m->params_bf16)
        if (p) cudaFree(p);
    for (float *p : m->params_x3)
        if (p) cudaFree(p);
    delete m;
    return SPK_OK;
}

extern "C" int64_t spk_model_add_param(spk_model_t *m, const float *host, int64_t n) {
    SPK_REQUIRE(m != nullptr && host != nullptr && n > 0, "bad parameter block");
    float *d = nullptr;
    SPK_CUDA_OK(cudaMalloc(&d, (size_t)n * sizeof(float)));
    cudaError_t e = cudaMemcpy(d, host, (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d);
        set_error("parameter upload failed: %s", cudaGetErrorString(e));
        return SPK_ERR_CUDA;
    }
    m->params.push_back(d);
    m->param_n.push_back(n);
    m->params_bf16.push_back(nullptr);
    m->params_x3.push_back(nullptr);
    return (int64_t)m->params.size() - 1;
}

extern "C" int spk_model_set_program(spk_model_t *m, int64_t T, const spk_buf_t *bufs, int32_t n_bufs,
                                     const spk_op_t *ops, int32_t n_ops) {
    SPK_REQUIRE(m != nullptr && bufs != nullptr && ops != nullptr, "null argument");
    SPK_REQUIRE(T > 0 && n_bufs >= 2 && n_ops > 0, "empty program");
    spk_program p;
    p.T = T;
    p.bufs.assign(bufs, bufs + n_bufs);
    p.ops.assign(ops, ops + n_ops);
    for (auto &b : p.bufs) SPK_REQUIRE(b.elems > 0 && (b.dtype == SPK_DT_F32 || b.dtype == SPK_DT_BF16), "bad buffer");
    SPK_REQUIRE(p.bufs[0].dtype == SPK_DT_F32 && p.bufs[1].dtype == SPK_DT_F32, "input/output buffers are f32");
    int rc = validate(m, p);
    if (rc != SPK_OK) return rc;
    if (m->precision == SPK_PREC_BF16) {
        // bf16 copies of every conv weight and prologue vector NOW, finished before the call returns: the forward
        // path then never launches a conversion kernel, so kernels started early by programmatic dependent launch
        // may stage their weights before waiting on the previous kernel
        const __nv_bfloat16 *tmp = nullptr;
        for (const spk_op_t &o : p.ops) {
            if (o.kind == SPK_OP_STEM_BLOCK) {
                rc = param_bf16(m, o.aux[0], &tmp, nullptr);
                if (rc == SPK_OK) rc = param_bf16(m, o.aux[3], &tmp, nullptr);
                if (rc != SPK_OK) return rc;
                continue;
            }
            if (o.kind != SPK_OP_CONV && o.kind != SPK_OP_CAM_LOCAL) continue;
            rc = param_bf16(m, o.w, &tmp, nullptr);
            if (rc == SPK_OK && o.pro_scale >= 0) rc = param_bf16(m, o.pro_scale, &tmp, nullptr);
            if (rc == SPK_OK && o.pro_shift >= 0) rc = param_bf16(m, o.pro_shift, &tmp, nullptr);
            if (rc != SPK_OK) return rc;
        }
        SPK_CUDA_OK(cudaDeviceSynchronize());
    }
    if (m->precision == SPK_PREC_F32) {
        // TF32 head / tail copies of the conv weights for the 3xTF32 tensor-core path, also finished here
        const float *tmp = nullptr;
        for (const spk_op_t &o : p.ops) {
            if (o.kind != SPK_OP_CONV && o.kind != SPK_OP_CAM_LOCAL) continue;
            rc = param_x3(m, o.w, &tmp, nullptr);
            if (rc != SPK_OK) return rc;
        }
        SPK_CUDA_OK(cudaDeviceSynchronize());
    }
    m->programs[T] = std::move(p);
    return SPK_OK;
}

extern "C" int64_t spk_model_workspace_bytes(spk_model_t *m, int64_t T, int64_t chunk, int64_t fine_chunk) {
    SPK_REQUIRE(m != nullptr && chunk > 0, "bad argument");
    auto it = m->programs.find(T);
    SPK_REQUIRE(it != m->programs.end(), "no program registered for T=%lld", (long long)T);
    if (fine_chunk <= 0 || fine_chunk > chunk) fine_chunk = chunk;
    std::vector<int64_t> off;
    std::vector<char> coarse;
    int64_t total = 0;
    plan(it->second, chunk, fine_chunk, off, coarse, total);
    return total;
}

extern "C" int spk_model_forward(spk_model_t *m, int64_t T, const float *feats, int64_t B, float *emb,
                                 void *workspace, int64_t workspace_bytes, int64_t chunk, int64_t fine_chunk,
                                 void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    SPK_REQUIRE(m != nullptr && feats != nullptr && emb != nullptr, "null argument");
    SPK_REQUIRE(B >= 0 && chunk > 0, "bad batch/chunk");
    auto it = m->programs.find(T);
    SPK_REQUIRE(it != m->programs.end(), "no program registered for T=%lld", (long long)T);
    const spk_program &p = it->second;
    if (fine_chunk <= 0 || fine_chunk > chunk) fine_chunk = chunk;
    std::vector<int64_t> off;
    std::vector<char> coarse;
    int64_t total = 0;
    plan(p, chunk, fine_chunk, off, coarse, total);
    if (total > workspace_bytes || (total > 0 && workspace == nullptr)) {
        set_error("workspace too small: need %lld bytes, got %lld", (long long)total, (long long)workspace_bytes);
        return SPK_ERR_WORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    NvtxRange whole("spk_model_forward");
    static const char *kOpNames[] = {"?", "stem", "conv", "cam_gate", "stats_pool", "aff_blend", "cam_local", "se_scale", "asp_pool", "stem_block"};

    bool has_phase1 = false;
    for (const spk_op_t &o : p.ops) has_phase1 |= (o.phase != 0);

    for (int64_t c0 = 0; c0 < B; c0 += chunk) {
      const int n_coarse = (int)std::min<int64_t>(chunk, B - c0);
      // pass 0: phase-0 ops per fine sub-batch; pass 1: phase-1 ops over the coarse sub-batch
      for (int pass = 0; pass < (has_phase1 ? 2 : 1); ++pass)
      for (int64_t f0 = 0; f0 < (pass == 0 ? n_coarse : 1); f0 += fine_chunk) {
        const int n = pass == 0 ? (int)std::min<int64_t>(fine_chunk, n_coarse - f0) : n_coarse;
        auto ptr = [&](int id) -> void * {
            if (id < 0) return nullptr;
            if (id == 0) return const_cast<float *>(feats) + (c0 + f0) * p.bufs[0].elems;
            if (id == 1) return emb + (c0 + f0) * p.bufs[1].elems;
            char *base = ws + off[id];
            if (coarse[id]) base += f0 * p.bufs[id].elems * (int64_t)dtype_size(p.bufs[id].dtype);
            return base;
        };
        auto dt = [&](int id) { return id < 0 ? SPK_DT_F32 : p.bufs[id].dtype; };
        for (size_t oi = 0; oi < p.ops.size(); ++oi) {
            const spk_op_t &o = p.ops[oi];
            if ((o.phase != 0) != (pass == 1)) continue;
            NvtxRange op_range(o.kind >= 1 && o.kind <= 9 ? kOpNames[o.kind] : kOpNames[0]);
            int rc = SPK_OK;
            switch (o.kind) {
                case SPK_OP_STEM: {
                    StemArgs a{};
                    a.feats = static_cast<const float *>(ptr(o.in_buf));
                    a.w = param(m, o.w);
                    a.scale = param(m, o.epi_scale);
                    a.shift = param(m, o.epi_shift);
                    a.y = ptr(o.out_buf);
                    a.B = n; a.T = o.W; a.F = o.H; a.Cout = o.Cout;
                    a.out_ld = o.out_ld; a.out_choff = o.out_choff; a.act = o.act;
                    rc = launch_stem(a, dt(o.out_buf), s);
                    break;
                }
                case SPK_OP_STEM_BLOCK: {
                    StemBlockArgs a{};
                    const __nv_bfloat16 *w1 = nullptr, *ws = nullptr;
                    rc = param_bf16(m, o.aux[0], &w1, s);
                    if (rc == SPK_OK) rc = param_bf16(m, o.aux[3], &ws, s);
                    if (rc != SPK_OK) break;
                    a.feats = static_cast<const float *>(ptr(o.in_buf));
                    a.w0 = param(m, o.w); a.s0 = param(m, o.epi_scale); a.b0 = param(m, o.epi_shift);
                    a.w1 = w1; a.s1 = param(m, o.aux[1]); a.b1 = param(m, o.aux[2]);
                    a.ws = ws; a.ss = param(m, o.iaux[0]); a.bs = param(m, o.iaux[1]);
                    a.y1 = ptr(o.out_buf); a.y2 = ptr(o.res_buf);
                    a.B = n; a.T = o.W; a.F = o.H; a.Ho = o.Ho;
                    a.y1_ld = o.out_ld; a.y1_choff = o.out_choff; a.y2_ld = o.res_ld; a.y2_choff = o.res_choff;
                    if (dt(o.out_buf) != SPK_DT_BF16 || dt(o.res_buf) != SPK_DT_BF16 || !stem_block_supported(a)) {
                        set_error("stem_block: unsupported shape or dtype (T=%d, F=%d)", a.T, a.F);
                        rc = SPK_ERR_UNSUPPORTED;
                        break;
                    }
                    rc = launch_stem_block(a, s);
                    break;
                }
                case SPK_OP_CONV:
                case SPK_OP_CAM_LOCAL: {
                    ConvArgs a{};
                    a.x = ptr(o.in_buf); a.y = ptr(o.out_buf); a.res = ptr(o.res_buf);
                    a.gate = static_cast<const float *>(ptr(o.gate_buf));
                    a.pro_scale = param(m, o.pro_scale); a.pro_shift = param(m, o.pro_shift);
                    a.epi_scale = param(m, o.epi_scale); a.epi_shift = param(m, o.epi_shift);
                    a.B = n; a.H = o.H; a.W = o.W; a.Cin = o.Cin; a.Ho = o.Ho; a.Wo = o.Wo; a.Cout = o.Cout;
                    a.KH = o.KH; a.KW = o.KW; a.sh = o.sh; a.sw = o.sw; a.ph = o.ph; a.pw = o.pw;
                    a.dh = o.dh; a.dw = o.dw;
                    a.in_ld = o.in_ld; a.in_choff = o.in_choff; a.out_ld = o.out_ld; a.out_choff = o.out_choff;
                    a.res_ld = o.res_ld; a.res_choff = o.res_choff;
                    a.gate_win = o.gate_win > 0 ? o.gate_win : 1;
                    a.gate_nwin = o.gate_buf >= 0 ? (o.Wo + a.gate_win - 1) / a.gate_win : 1;
                    a.pro_relu = o.pro_relu; a.act = o.act;
                    a.K = o.KH * o.KW * o.Cin;
                    a.M = (long long)n * o.Ho * o.Wo;
                    a.l2_flags = o.reserved;
                    if (m->precision == SPK_PREC_BF16 && o.pro_scale >= 0) {      // converted at set_program time
                        const __nv_bfloat16 *psb = nullptr, *phb = nullptr;
                        rc = param_bf16(m, o.pro_scale, &psb, s);
                        if (rc == SPK_OK) rc = param_bf16(m, o.pro_shift, &phb, s);
                        if (rc != SPK_OK) break;
                        a.pro_scale_bf = psb; a.pro_shift_bf = phb;
                    }
                    if (o.kind == SPK_OP_CONV) {            // ECAPA extras (see b200spk.h)
                        a.post_scale = param(m, o.aux[0]); a.post_shift = param(m, o.aux[1]);
                        a.post_act = o.iaux[0]; a.pad_reflect = o.iaux[1]; a.gate_additive = o.iaux[2];
                    }
                    if (o.kind == SPK_OP_CAM_LOCAL) {
                        const int hidden = o.iaux[0], seg_len = o.iaux[1];
                        a.gate_win = seg_len;
                        a.gate_nwin = (o.Wo + seg_len - 1) / seg_len;
                        ConvArgs f = a;
                        f.gate = nullptr;
                        if (m->precision == SPK_PREC_BF16 &&
                            cam_local_supported(f, dt(o.in_buf), dt(o.out_buf), hidden, seg_len)) {
                            const __nv_bfloat16 *wb = nullptr;
                            rc = param_bf16(m, o.w, &wb, s);
                            if (rc == SPK_OK) {
                                f.w = wb;
                                rc = launch_cam_local(f, param(m, o.iaux[2]), param(m, o.aux[1]), param(m, o.iaux[3]),
                                                      param(m, o.aux[3]), hidden, seg_len, s);
                            }
                            break;
                        }
                        // unfused: context gate into the scratch buffer, then the gated conv below
                        CamGateArgs cg{};
                        cg.x = ptr(o.in_buf);
                        cg.gate = static_cast<float *>(ptr(o.gate_buf));
                        cg.w1 = param(m, o.aux[0]); cg.b1 = param(m, o.aux[1]);
                        cg.w2 = param(m, o.aux[2]); cg.b2 = param(m, o.aux[3]);
                        cg.B = n; cg.T = o.W; cg.C = o.Cin; cg.in_ld = o.in_ld; cg.in_choff = o.in_choff;
                        cg.hidden = hidden; cg.seg_len = seg_len; cg.Cout = o.Cout; cg.nwin = a.gate_nwin;
                        rc = launch_cam_gate(cg, dt(o.in_buf), s);
                        if (rc != SPK_OK) break;
                    }
                    if (m->precision == SPK_PREC_BF16 &&
                        conv_slab3_supported(a, dt(o.in_buf), dt(o.out_buf), dt(o.res_buf))) {
                        const __nv_bfloat16 *wb = nullptr;
                        rc = param_bf16(m, o.w, &wb, s);
                        if (rc == SPK_OK) {
                            a.w = wb;
                            trace_dispatch("slab3", a);
                            rc = launch_conv_slab3(a, s);
                        }
                    } else if (m->precision == SPK_PREC_BF16 &&
                        conv_slab4_supported(a, dt(o.in_buf), dt(o.out_buf), dt(o.res_buf))) {
                        const __nv_bfloat16 *wb = nullptr;
                        rc = param_bf16(m, o.w, &wb, s);
                        if (rc == SPK_OK) {
                            a.w = wb;
                            trace_dispatch("slab4", a);
                            rc = launch_conv_slab4(a, s);
                        }
                    } else if (m->precision == SPK_PREC_BF16 &&
                        conv_slab_supported(a, dt(o.in_buf), dt(o.out_buf), dt(o.res_buf))) {
                        const __nv_bfloat16 *wb = nullptr;
                        rc = param_bf16(m, o.w, &wb, s);
                        if (rc == SPK_OK) {
                            a.w = wb;
                            trace_dispatch("slab2", a);
                            rc = launch_conv_slab(a, s);
                        }
                    } else if (m->precision == SPK_PREC_BF16 && conv_gemm_supported(a, dt(o.in_buf))) {
                        const __nv_bfloat16 *wb = nullptr;
                        rc = param_bf16(m, o.w, &wb, s);
                        if (rc == SPK_OK) {
                            a.w = wb;
                            trace_dispatch(a.KH * a.KW == 1 && a.sh == 1 && a.sw == 1 ? "gemm" : "gemm_im2col", a);
                            rc = launch_conv_gemm(a, dt(o.out_buf), dt(o.res_buf), s);
                        }
                    } else if (m->precision == SPK_PREC_BF16 && conv_tc_supported(a, dt(o.in_buf))) {
                        const __nv_bfloat16 *wb = nullptr;
                        rc = param_bf16(m, o.w, &wb, s);
                        if (rc == SPK_OK) {
                            a.w = wb;
                            trace_dispatch("tc2_gather", a);
                            rc = launch_conv_tc(a, dt(o.out_buf), dt(o.res_buf), s);
                        }
                    } else if (m->precision == SPK_PREC_F32 && conv_f32x3_supported(a, dt(o.in_buf), dt(o.out_buf), dt(o.res_buf))) {
                        const float *w3 = nullptr;
                        rc = param_x3(m, o.w, &w3, s);
                        if (rc == SPK_OK) {
                            a.w = w3;
                            trace_dispatch("f32x3", a);
                            rc = launch_conv_f32x3(a, s);
                            if (rc == SPK_OK && a.pad_reflect && a.M > 0) {      // mirrored taps at the segment ends, fp32 weights
                                ConvArgs f = a;
                                f.w = param(m, o.w);
                                rc = launch_reflect_edge_fix(f, SPK_DT_F32, SPK_DT_F32, s);
                            }
                        }
                    } else {
                        a.w = param(m, o.w);
                        trace_dispatch("simt", a);
                        rc = launch_conv_simt(a, dt(o.in_buf), dt(o.out_buf), dt(o.res_buf), s);
                    }
                    break;
                }
                case SPK_OP_CAM_GATE: {
                    CamGateArgs a{};
                    a.x = ptr(o.in_buf);
                    a.gate = static_cast<float *>(ptr(o.out_buf));
                    a.w1 = param(m, o.aux[0]); a.b1 = param(m, o.aux[1]);
                    a.w2 = param(m, o.aux[2]); a.b2 = param(m, o.aux[3]);
                    a.B = n; a.T = o.W; a.C = o.Cin; a.in_ld = o.in_ld; a.in_choff = o.in_choff;
                    a.hidden = o.iaux[0]; a.seg_len = o.iaux[1]; a.Cout = o.Cout;
                    a.nwin = (o.W + a.seg_len - 1) / a.seg_len;
                    a.se_mode = o.iaux[2];
                    rc = launch_cam_gate(a, dt(o.in_buf), s);
                    break;
                }
                case SPK_OP_STATS_POOL: {
                    StatsPoolArgs a{};
                    a.x = ptr(o.in_buf);
                    a.y = static_cast<float *>(ptr(o.out_buf));
                    a.B = n; a.G = o.H; a.P = o.W; a.C = o.Cin; a.in_ld = o.in_ld; a.in_choff = o.in_choff;
                    a.unbiased = o.iaux[0]; a.eps = o.faux[0]; a.var_floor = o.faux[1];
                    rc = launch_stats_pool(a, dt(o.in_buf), s);
                    break;
                }
                case SPK_OP_AFF_BLEND: {
                    AffBlendArgs a{};
                    a.x = ptr(o.in_buf); a.y = ptr(o.res_buf); a.z = ptr(o.gate_buf); a.out = ptr(o.out_buf);
                    a.M = (long long)n * o.H * o.W; a.C = o.Cin;
                    a.x_ld = o.in_ld; a.x_choff = o.in_choff; a.y_ld = o.res_ld; a.y_choff = o.res_choff;
                    a.z_ld = o.iaux[0]; a.z_choff = o.iaux[1]; a.out_ld = o.out_ld; a.out_choff = o.out_choff;
                    rc = launch_aff_blend(a, dt(o.in_buf), dt(o.out_buf), s);
                    break;
                }
                case SPK_OP_SE_SCALE: {
                    SeScaleArgs a{};
                    a.x = ptr(o.in_buf); a.res = ptr(o.res_buf); a.gate = static_cast<const float *>(ptr(o.gate_buf)); a.out = ptr(o.out_buf);
                    a.B = n; a.P = o.H * o.W; a.C = o.Cin;
                    a.x_ld = o.in_ld; a.x_choff = o.in_choff; a.res_ld = o.res_ld; a.res_choff = o.res_choff;
                    a.out_ld = o.out_ld; a.out_choff = o.out_choff;
                    rc = launch_se_scale(a, dt(o.in_buf), dt(o.res_buf), dt(o.out_buf), s);
                    break;
                }
                case SPK_OP_ASP_POOL: {
                    AspPoolArgs a{};
                    a.logits = ptr(o.in_buf); a.x = ptr(o.res_buf); a.out = static_cast<float *>(ptr(o.out_buf));
                    a.B = n; a.P = o.H * o.W; a.C = o.Cin;
                    a.l_ld = o.in_ld; a.l_choff = o.in_choff; a.x_ld = o.res_ld; a.x_choff = o.res_choff;
                    a.var_floor = o.faux[0];
                    rc = launch_asp_pool(a, dt(o.in_buf), dt(o.res_buf), s);
                    break;
                }
                default:
                    set_error("unknown op kind %d", o.kind);
                    rc = SPK_ERR_INVALID;
            }
            if (rc != SPK_OK) return rc;
        }
      }
    }
    return SPK_OK;
}

extern "C" int spk_model_read_buffer(spk_model_t *m, int64_t T, int32_t buf_id, int64_t chunk, int64_t fine_chunk,
                                     const void *workspace, float *dst, int64_t n, void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    SPK_REQUIRE(m != nullptr && workspace != nullptr && dst != nullptr, "null argument");
    auto it = m->programs.find(T);
    SPK_REQUIRE(it != m->programs.end(), "no program registered for T=%lld", (long long)T);
    const spk_program &p = it->second;
    SPK_REQUIRE(buf_id >= 2 && buf_id < (int)p.bufs.size(), "buffer id %d is not a workspace buffer", buf_id);
    if (fine_chunk <= 0 || fine_chunk > chunk) fine_chunk = chunk;
    std::vector<int64_t> off;
    std::vector<char> coarse;
    int64_t total = 0;
    plan(p, chunk, fine_chunk, off, coarse, total);
    SPK_REQUIRE(n <= p.bufs[buf_id].elems * (coarse[buf_id] ? chunk : fine_chunk), "read past the end of buffer %d",
                buf_id);
    return launch_widen(static_cast<const char *>(workspace) + off[buf_id], p.bufs[buf_id].dtype, dst, n, s);
}
