"""Drop-in for ``speakerlab.models.ecapa_tdnn.ECAPA_TDNN.ECAPA_TDNN`` (ECAPA_TDNN.py:356-463).

Same constructor signature and ``state_dict`` layout (231 keys for ``channels=[1024]*4+[3072]``).
``forward(x[B, T, F])`` (``lengths`` must be None: whole segments, as every reference call site
uses it) compiles the eval-mode network into the fused-op list of ``libb200spk``.  Inference only;
``groups`` must be all ones, ``activation`` ReLU, ``global_context`` True (the shipped config).

Op mapping (reference file:line -> fused op); every TDNNBlock is conv(+bias) -> ReLU -> BatchNorm
(:150-151), i.e. a CONV with the bias as epilogue shift and the folded BN as the post-activation
affine; 'same' reflect padding (:42-106) is handled inside the conv gather:
  blocks[0] TDNN(F -> C, k=5)                       :399-408   (bf16 mode: AFF_BLEND copy casts the features first) CONV
  SERes2NetBlock.tdnn1 / tdnn2 (1x1)                :310-326   CONV
  Res2NetBlock: y0 = x0; y1 = blk(x1); yi = blk(xi + y(i-1))   :180-191   AFF_BLEND copy / AFF_BLEND add + CONV (k=3, dilated) writing channel slices
  SEBlock: mean_T -> 1x1 -> ReLU -> 1x1 -> sigmoid  :209-222   CAM_GATE in squeeze-excitation mode
  s * x + residual                                  :222,345   SE_SCALE, written straight into its slice of the MFA concat buffer (:447)
  mfa TDNN(3C -> 3C, k=1)                           :424-431   CONV
  ASP global context: mean, sqrt(clamp(var, 1e-12)) :253-272   STATS_POOL (biased, variance floor)
  ASP tdnn on cat([x, mean, std]) (9216 -> 128)     :236,277   the mean/std columns are constant over T: a per-segment bias
                                                               u = W[:, C:3C] [mean; std] + b (tiny CONV on the stats) added before
                                                               the ReLU of CONV(W[:, :C]) -> BN -> tanh (post affine + post act)
  ASP conv (128 -> 3C), softmax over T, weighted mean/std     :277-285   CONV, ASP_POOL
  asp_bn + fc (6144 -> lin_neurons, bias)           :434-441   CONV with the BN as prologue
"""
import torch
from torch import nn

from . import _lib
from .program import EngineBase, EngineModule, Program, fold_bn


class _Conv1d(nn.Module):
    def __init__(self, out_channels, kernel_size, in_channels, dilation=1):
        super().__init__()
        self.kernel_size, self.dilation = kernel_size, dilation
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, dilation=dilation, padding=0, bias=True)


class _BatchNorm1d(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.norm = nn.BatchNorm1d(input_size)


class _TDNNBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, dilation):
        super().__init__()
        self.conv = _Conv1d(out_channels, kernel_size, in_channels, dilation)
        self.activation = nn.ReLU()
        self.norm = _BatchNorm1d(out_channels)


class _Res2NetBlock(nn.Module):
    def __init__(self, in_channels, out_channels, scale, kernel_size, dilation):
        super().__init__()
        assert in_channels % scale == 0 and out_channels % scale == 0
        self.blocks = nn.ModuleList([_TDNNBlock(in_channels // scale, out_channels // scale, kernel_size, dilation)
                                     for _ in range(scale - 1)])
        self.scale = scale


class _SEBlock(nn.Module):
    def __init__(self, in_channels, se_channels, out_channels):
        super().__init__()
        self.conv1 = _Conv1d(se_channels, 1, in_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _Conv1d(out_channels, 1, se_channels)
        self.sigmoid = nn.Sigmoid()


class _SERes2NetBlock(nn.Module):
    def __init__(self, in_channels, out_channels, res2net_scale, se_channels, kernel_size, dilation):
        super().__init__()
        self.out_channels = out_channels
        self.tdnn1 = _TDNNBlock(in_channels, out_channels, 1, 1)
        self.res2net_block = _Res2NetBlock(out_channels, out_channels, res2net_scale, kernel_size, dilation)
        self.tdnn2 = _TDNNBlock(out_channels, out_channels, 1, 1)
        self.se_block = _SEBlock(out_channels, se_channels, out_channels)
        self.shortcut = None
        if in_channels != out_channels:
            self.shortcut = _Conv1d(out_channels, 1, in_channels)


class _ASP(nn.Module):
    def __init__(self, channels, attention_channels):
        super().__init__()
        self.eps = 1e-12
        self.tdnn = _TDNNBlock(channels * 3, attention_channels, 1, 1)
        self.tanh = nn.Tanh()
        self.conv = _Conv1d(channels, 1, attention_channels)


class ECAPA_TDNN(EngineModule):
    def __init__(self, input_size, device="cpu", lin_neurons=192, activation=torch.nn.ReLU,
                 channels=[512, 512, 512, 512, 1536], kernel_sizes=[5, 3, 3, 3, 1], dilations=[1, 2, 3, 4, 1],
                 attention_channels=128, res2net_scale=8, se_channels=128, global_context=True, groups=[1, 1, 1, 1, 1],
                 precision="fp32", chunk=None):
        super().__init__()
        assert len(channels) == len(kernel_sizes)
        assert len(channels) == len(dilations)
        assert activation is torch.nn.ReLU and global_context and all(g == 1 for g in groups), \
            "only the shipped configuration (ReLU, global context, no grouped convs) is implemented"
        assert input_size % 8 == 0 and all(c % (8 * res2net_scale) == 0 for c in channels[:-1]) and channels[-1] % 16 == 0
        assert channels[-1] == sum(channels[1:-1]), "mfa consumes the concatenated SE-Res2Net outputs"
        self._init_engine(precision, chunk)
        self.input_size, self.lin_neurons = input_size, lin_neurons
        self.channels, self.kernel_sizes, self.dilations = list(channels), list(kernel_sizes), list(dilations)
        self.attention_channels, self.res2net_scale, self.se_channels = attention_channels, res2net_scale, se_channels
        self.blocks = nn.ModuleList()
        self.blocks.append(_TDNNBlock(input_size, channels[0], kernel_sizes[0], dilations[0]))
        for i in range(1, len(channels) - 1):
            self.blocks.append(_SERes2NetBlock(channels[i - 1], channels[i], res2net_scale, se_channels, kernel_sizes[i],
                                               dilations[i]))
        self.mfa = _TDNNBlock(channels[-1], channels[-1], kernel_sizes[-1], dilations[-1])
        self.asp = _ASP(channels[-1], attention_channels)
        self.asp_bn = _BatchNorm1d(channels[-1] * 2)
        self.fc = _Conv1d(lin_neurons, 1, channels[-1] * 2)
        self.eval()

    def forward(self, x, lengths=None):
        """x [B, T, input_size] -> [B, lin_neurons] (float32, same device)."""
        assert lengths is None, "relative lengths are not supported: pass whole segments"
        return self._run(x, self.input_size, self.lin_neurons)


class _Engine(EngineBase):
    def default_chunks(self, T):
        mod = self.m
        width = mod.channels[0] + 5 * max(mod.channels[1:-1]) + 3 * mod.channels[-1]
        per_seg = T * width * (2 if self.model.precision == _lib.PREC_BF16 else 4)
        c = max(1, min(512, int(6e9 // per_seg)))
        return c, c

    # ---- parameter helpers
    def _ones(self, n):
        return self._p(("ones", n), lambda: torch.ones(n))

    def _w1d_cols(self, key, c0, c1):
        """[Cout, Cin, 1] conv weight restricted to input channels [c0, c1) -> packed [Cout, 1, 1, c1-c0]."""
        return self._p(("wcols", key, c0, c1), lambda: self.sd[key][:, c0:c1, 0].contiguous())

    def _tdnn(self, prog, prefix, *, in_buf, in_ld, in_choff, Cin, out_buf, out_ld, out_choff, Cout, T, k, dil, w=None, **extra):
        """One TDNNBlock: reflect-padded conv + bias -> ReLU -> BN."""
        ps, pb = self._bn(prefix + ".norm.norm")
        prog.op(_lib.OP_CONV, in_buf=in_buf, in_ld=in_ld, in_choff=in_choff, out_buf=out_buf, out_ld=out_ld, out_choff=out_choff,
                H=1, W=T, Cin=Cin, Ho=1, Wo=T, Cout=Cout, KH=1, KW=k, pw=dil * (k - 1) // 2, dw=dil,
                w=self._w1d(prefix + ".conv.conv.weight") if w is None else w,
                epi_scale=self._ones(Cout), epi_shift=self._raw(prefix + ".conv.conv.bias"), act=_lib.ACT_RELU,
                aux=[ps, pb, -1, -1], iaux=[extra.pop("post_act", _lib.ACT_NONE), 1 if k > 1 else 0, extra.pop("additive", 0), 0], **extra)

    def compile(self, T):
        mod, AD = self.m, self.model.act_dtype
        F, E, A = mod.input_size, mod.lin_neurons, mod.attention_channels
        ch, ks, dl = mod.channels, mod.kernel_sizes, mod.dilations
        Cm = ch[-1]
        nblk = len(ch) - 2
        assert T > max(d * (k - 1) // 2 for k, d in zip(ks, dl)), "segment shorter than the reflect padding"
        prog = Program(T * F, E)
        # ---- blocks[0]
        x0 = 0
        if AD != _lib.DT_F32:
            x0 = prog.buf("x0", T * F, AD)
            prog.op(_lib.OP_AFF_BLEND, in_buf=0, in_ld=F, out_buf=x0, out_ld=F, H=1, W=T, Cin=F)
        b0 = prog.buf("block0", T * ch[0], AD)
        self._tdnn(prog, "blocks.0", in_buf=x0, in_ld=F, in_choff=0, Cin=F, out_buf=b0, out_ld=ch[0], out_choff=0, Cout=ch[0],
                   T=T, k=ks[0], dil=dl[0])
        cat = prog.buf("cat", T * Cm, AD)
        cur = (b0, ch[0], 0, ch[0])            # buffer, ld, choff, channels
        off = 0
        for i in range(1, nblk + 1):
            C, pre = ch[i], "blocks.%d" % i
            S = mod.res2net_scale
            w = C // S
            in_buf, in_ld, in_choff, Cin = cur
            res = (in_buf, in_ld, in_choff)
            if Cin != C:
                sc = prog.buf(pre + ".shortcut", T * C, AD)
                prog.op(_lib.OP_CONV, in_buf=in_buf, in_ld=in_ld, in_choff=in_choff, out_buf=sc, out_ld=C, H=1, W=T, Cin=Cin, Ho=1,
                        Wo=T, Cout=C, w=self._w1d(pre + ".shortcut.conv.weight"), epi_scale=self._ones(C),
                        epi_shift=self._raw(pre + ".shortcut.conv.bias"))
                res = (sc, C, 0)
            t1 = prog.buf(pre + ".tdnn1", T * C, AD)
            self._tdnn(prog, pre + ".tdnn1", in_buf=in_buf, in_ld=in_ld, in_choff=in_choff, Cin=Cin, out_buf=t1, out_ld=C,
                       out_choff=0, Cout=C, T=T, k=1, dil=1)
            r = prog.buf(pre + ".res2net", T * C, AD)
            tmp = prog.buf(pre + ".sum", T * w, AD)
            prog.op(_lib.OP_AFF_BLEND, in_buf=t1, in_ld=C, in_choff=0, out_buf=r, out_ld=C, out_choff=0, H=1, W=T, Cin=w)
            for j in range(1, S):
                src = (t1, C, j * w)
                if j >= 2:
                    prog.op(_lib.OP_AFF_BLEND, in_buf=t1, in_ld=C, in_choff=j * w, res_buf=r, res_ld=C, res_choff=(j - 1) * w,
                            out_buf=tmp, out_ld=w, out_choff=0, H=1, W=T, Cin=w)
                    src = (tmp, w, 0)
                self._tdnn(prog, "%s.res2net_block.blocks.%d" % (pre, j - 1), in_buf=src[0], in_ld=src[1], in_choff=src[2], Cin=w,
                           out_buf=r, out_ld=C, out_choff=j * w, Cout=w, T=T, k=ks[i], dil=dl[i])
            t2 = prog.buf(pre + ".tdnn2", T * C, AD)
            self._tdnn(prog, pre + ".tdnn2", in_buf=r, in_ld=C, in_choff=0, Cin=C, out_buf=t2, out_ld=C, out_choff=0, Cout=C, T=T,
                       k=1, dil=1)
            gate = prog.buf(pre + ".se", C, _lib.DT_F32)
            se = pre + ".se_block"
            prog.op(_lib.OP_CAM_GATE, in_buf=t2, in_ld=C, out_buf=gate, W=T, Cin=C, Cout=C,
                    aux=[self._raw(se + ".conv1.conv.weight"), self._raw(se + ".conv1.conv.bias"),
                         self._raw(se + ".conv2.conv.weight"), self._raw(se + ".conv2.conv.bias")],
                    iaux=[mod.se_channels, T, 1, 0])
            prog.op(_lib.OP_SE_SCALE, in_buf=t2, in_ld=C, res_buf=res[0], res_ld=res[1], res_choff=res[2], gate_buf=gate,
                    out_buf=cat, out_ld=Cm, out_choff=off, H=1, W=T, Cin=C)
            cur = (cat, Cm, off, C)
            off += C
        # ---- multi-layer feature aggregation
        m = prog.buf("mfa", T * Cm, AD)
        self._tdnn(prog, "mfa", in_buf=cat, in_ld=Cm, in_choff=0, Cin=Cm, out_buf=m, out_ld=Cm, out_choff=0, Cout=Cm, T=T,
                   k=ks[-1], dil=dl[-1])
        # ---- attentive statistics pooling
        stats = prog.buf("asp.stats", 2 * Cm, _lib.DT_F32)
        prog.op(_lib.OP_STATS_POOL, in_buf=m, in_ld=Cm, out_buf=stats, H=1, W=T, Cin=Cm, iaux=[0, 0, 0, 0], faux=[0.0, 1e-12])
        u = prog.buf("asp.ctx_bias", A, _lib.DT_F32)
        wkey = "asp.tdnn.conv.conv.weight"
        prog.op(_lib.OP_CONV, in_buf=stats, in_ld=2 * Cm, out_buf=u, out_ld=A, H=1, W=1, Cin=2 * Cm, Ho=1, Wo=1, Cout=A,
                w=self._w1d_cols(wkey, Cm, 3 * Cm), epi_scale=self._ones(A), epi_shift=self._raw("asp.tdnn.conv.conv.bias"))
        att = prog.buf("asp.att", T * A, AD)
        ps, pb = self._bn("asp.tdnn.norm.norm")
        prog.op(_lib.OP_CONV, in_buf=m, in_ld=Cm, out_buf=att, out_ld=A, H=1, W=T, Cin=Cm, Ho=1, Wo=T, Cout=A,
                w=self._w1d_cols(wkey, 0, Cm), gate_buf=u, gate_win=T, act=_lib.ACT_RELU, aux=[ps, pb, -1, -1],
                iaux=[_lib.ACT_TANH, 0, 1, 0])
        logits = prog.buf("asp.logits", T * Cm, AD)
        prog.op(_lib.OP_CONV, in_buf=att, in_ld=A, out_buf=logits, out_ld=Cm, H=1, W=T, Cin=A, Ho=1, Wo=T, Cout=Cm,
                w=self._w1d("asp.conv.conv.weight"), epi_scale=self._ones(Cm), epi_shift=self._raw("asp.conv.conv.bias"))
        pooled = prog.buf("asp.pooled", 2 * Cm, _lib.DT_F32)
        prog.op(_lib.OP_ASP_POOL, in_buf=logits, in_ld=Cm, res_buf=m, res_ld=Cm, out_buf=pooled, H=1, W=T, Cin=Cm, faux=[1e-12, 0.0])
        # ---- asp_bn + fc
        bs, bb = self._bn("asp_bn.norm")
        prog.op(_lib.OP_CONV, in_buf=pooled, in_ld=2 * Cm, out_buf=1, out_ld=E, H=1, W=1, Cin=2 * Cm, Ho=1, Wo=1, Cout=E,
                w=self._w1d("fc.conv.weight"), pro_scale=bs, pro_shift=bb, pro_relu=0, epi_scale=self._ones(E),
                epi_shift=self._raw("fc.conv.bias"))
        self.model.set_program(T, prog)


ECAPA_TDNN.engine_cls = _Engine
