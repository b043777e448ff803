"""Drop-in for ``speakerlab.models.campplus.DTDNN.CAMPPlus`` (DTDNN.py:50-115).

Same constructor signature and the same ``state_dict`` layout (937 keys for the default
configuration), so reference checkpoints load unchanged; ``forward`` does not run PyTorch ops:
it compiles the eval-mode network into the fused-op list of ``libb200spk`` (BN folded, weights
repacked channels-last) and runs it with the sm_100a kernels.  Inference only.

Op mapping (reference file:line -> fused op):
  FCM head conv1+bn1+relu            DTDNN.py:40-41        STEM
  BasicResBlock                      layers.py:248-253     CONV(+bn,relu) / CONV(shortcut+bn) / CONV(+bn,+res,relu)
  head conv2+bn2+relu, reshape       DTDNN.py:44-48        CONV; reshape is free (tdnn reads [10,T,32] as K=(f,k,c))
  TDNNLayer k5 s2 +bn+relu           layers.py:40-67       CONV KH=10,KW=5 writing channels [0,128) of block1's buffer
  CAMDenseTDNNLayer.bn_function      layers.py:140-141     CONV 1x1, BN-ReLU prologue on the growing concat, BN-ReLU epilogue
  CAMLayer (context gate, local     layers.py:93-99,179   CAM_LOCAL: gate evaluated once per distinct 100-frame window,
  conv, gating) + torch.cat                                dilated k3 conv x gate written in place into the concat buffer
  TransitLayer                       layers.py:193-196     CONV 1x1 with BN-ReLU prologue (transit3 also folds out_nonlinear)
  StatsPool                          layers.py:26-37       STATS_POOL (unbiased std)
  DenseLayer + BatchNorm1d(affine=F) layers.py:209-215     CONV 1x1 on [1024] with folded BN epilogue
"""
from collections import OrderedDict

import os

import torch
from torch import nn

from . import _lib
from .program import EngineBase, EngineModule, Program, conv_out

SEG_LEN = 100   # CAMLayer.seg_pooling (layers.py:100)


class _Tree(nn.Module):
    """Parameter container with named children (holds weights; never runs a PyTorch forward)."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.add_module(k, v)


def _nonlinear(channels, affine=True):
    return _Tree(batchnorm=nn.BatchNorm1d(channels, affine=affine))


def _res_block(cin, cout, stride):
    t = _Tree(conv1=nn.Conv2d(cin, cout, 3, stride=(stride, 1), padding=1, bias=False), bn1=nn.BatchNorm2d(cout),
              conv2=nn.Conv2d(cout, cout, 3, padding=1, bias=False), bn2=nn.BatchNorm2d(cout))
    sc = nn.Sequential()
    if stride != 1 or cin != cout:
        sc = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=(stride, 1), bias=False), nn.BatchNorm2d(cout))
    t.add_module("shortcut", sc)
    return t


class CAMPPlus(EngineModule):
    def __init__(self, feat_dim=80, embedding_size=512, growth_rate=32, bn_size=4, init_channels=128,
                 config_str='batchnorm-relu', memory_efficient=True, precision="fp32", chunk=None):
        super().__init__()
        assert config_str == 'batchnorm-relu', "only the shipped 'batchnorm-relu' configuration is implemented"
        assert feat_dim % 8 == 0
        self.feat_dim, self.embedding_size = feat_dim, embedding_size
        self.growth_rate, self.bn_channels, self.init_channels = growth_rate, bn_size * growth_rate, init_channels
        self.block_cfg = ((12, 3, 1), (24, 3, 2), (16, 3, 2))   # DTDNN.py:77-78
        self._init_engine(precision, chunk)
        m = 32
        self.head = _Tree(conv1=nn.Conv2d(1, m, 3, padding=1, bias=False), bn1=nn.BatchNorm2d(m),
                          layer1=nn.Sequential(_res_block(m, m, 2), _res_block(m, m, 1)),
                          layer2=nn.Sequential(_res_block(m, m, 2), _res_block(m, m, 1)),
                          conv2=nn.Conv2d(m, m, 3, stride=(2, 1), padding=1, bias=False), bn2=nn.BatchNorm2d(m))
        channels = m * (feat_dim // 8)
        xv = OrderedDict()
        xv["tdnn"] = _Tree(linear=nn.Conv1d(channels, init_channels, 5, stride=2, padding=2, bias=False),
                           nonlinear=_nonlinear(init_channels))
        channels = init_channels
        for i, (n_layers, k, dil) in enumerate(self.block_cfg, start=1):
            blk = nn.ModuleList()
            for j in range(n_layers):
                cin = channels + j * growth_rate
                layer = _Tree(nonlinear1=_nonlinear(cin), linear1=nn.Conv1d(cin, self.bn_channels, 1, bias=False),
                              nonlinear2=_nonlinear(self.bn_channels),
                              cam_layer=_Tree(
                                  linear_local=nn.Conv1d(self.bn_channels, growth_rate, k, padding=dil, dilation=dil,
                                                         bias=False),
                                  linear1=nn.Conv1d(self.bn_channels, self.bn_channels // 2, 1),
                                  linear2=nn.Conv1d(self.bn_channels // 2, growth_rate, 1)))
                blk.add_module("tdnnd%d" % (j + 1), layer)
            xv["block%d" % i] = blk
            channels += n_layers * growth_rate
            xv["transit%d" % i] = _Tree(nonlinear=_nonlinear(channels),
                                        linear=nn.Conv1d(channels, channels // 2, 1, bias=False))
            channels //= 2
        xv["out_nonlinear"] = _nonlinear(channels)
        xv["dense"] = _Tree(linear=nn.Conv1d(channels * 2, embedding_size, 1, bias=False),
                            nonlinear=_nonlinear(embedding_size, affine=False))
        self.xvector = _Tree(**xv)
        self.final_channels = channels
        for mod in self.modules():          # DTDNN.py:105-109
            if isinstance(mod, (nn.Conv1d, nn.Linear)):
                nn.init.kaiming_normal_(mod.weight.data)
                if mod.bias is not None:
                    nn.init.zeros_(mod.bias)
        self.eval()

    def forward(self, x):
        """x [B, T, feat_dim] -> [B, embedding_size] (float32, same device)."""
        return self._run(x, self.feat_dim, self.embedding_size)


L2_HINTS = int(os.environ.get("SPK_L2_HINTS", "3"))       # bit 0: stream the concat reads, bit 1: keep the bottleneck output in L2


class _Engine(EngineBase):
    def default_chunks(self, T):
        """(coarse, fine) sub-batch sizes.  Measured on B200 (tools/sweep_chunks.py): launch
        count and pipeline fill matter more than L2 residency of the 2-D front - every kernel of the D-TDNN part
        pays ~10 us of ramp and tail, ~20 % of a 2048-segment launch - so both are large (16,384 windows: 2048/1024
        255 k emb/s, 8192/2048 275 k, 16384/4096 279 k); they only bound the workspace (about 0.6 MB coarse + 1.9 MB
        fine per segment in bf16: 8.8 GB at 8192/2048, allocated for the batch actually seen)."""
        scale = max(1, T // 148)
        return max(64, 8192 // scale), max(32, 2048 // scale)

    def compile(self, T):
        mod, AD = self.m, self.model.act_dtype
        F, E = mod.feat_dim, mod.embedding_size
        C0 = 32
        prog = Program(T * F, E)
        big = prog.buf("fcm_a", F * T * C0, AD)
        h1 = F // 2
        bb = prog.buf("fcm_b", h1 * T * C0, AD)
        bc = prog.buf("fcm_c", h1 * T * C0, AD)
        bd = prog.buf("fcm_d", h1 * T * C0, AD)
        # stem.  In bf16 mode, for segments up to 254 frames, the stem is fused with layer1[0]'s conv1 and shortcut
        # (SPK_OP_STEM_BLOCK): its output, the largest tensor of the network, is then never stored.
        fuse_stem = (self.model.precision == _lib.PREC_BF16 and T <= 254 and F % 2 == 0 and T >= 8 and
                     os.environ.get("SPK_NO_STEM_FUSE", "0") != "1" and "head.layer1.0.shortcut.0.weight" in self.sd)
        s, b = self._bn("head.bn1")
        if not fuse_stem:
            prog.op(_lib.OP_STEM, in_buf=0, out_buf=big, out_ld=C0, H=F, W=T, Cout=C0,
                    w=self._raw("head.conv1.weight"), epi_scale=s, epi_shift=b, act=_lib.ACT_RELU)

        def conv3x3(src, dst, H, stride, wkey, bnkey, act, res=-1):
            Ho = conv_out(H, 3, stride, 1)
            sc, sh = self._bn(bnkey)
            prog.op(_lib.OP_CONV, in_buf=src, in_ld=C0, out_buf=dst, out_ld=C0, res_buf=res, res_ld=C0,
                    H=H, W=T, Cin=C0, Ho=Ho, Wo=T, Cout=C0, KH=3, KW=3, sh=stride, sw=1, ph=1, pw=1,
                    w=self._w2d(wkey), epi_scale=sc, epi_shift=sh, act=act)
            return Ho

        def res_block(prefix, src, H, stride, tmp, sc_buf, dst):
            """dst = relu(bn2(conv2(relu(bn1(conv1(src))))) + shortcut(src))"""
            Ho = conv3x3(src, tmp, H, stride, prefix + ".conv1.weight", prefix + ".bn1", _lib.ACT_RELU)
            if (prefix + ".shortcut.0.weight") in self.sd:
                sc, sh = self._bn(prefix + ".shortcut.1")
                prog.op(_lib.OP_CONV, in_buf=src, in_ld=C0, out_buf=sc_buf, out_ld=C0, H=H, W=T, Cin=C0, Ho=Ho, Wo=T,
                        Cout=C0, KH=1, KW=1, sh=stride, sw=1, w=self._w2d(prefix + ".shortcut.0.weight"),
                        epi_scale=sc, epi_shift=sh, act=_lib.ACT_NONE)
                res = sc_buf
            else:
                res = src
            conv3x3(tmp, dst, Ho, 1, prefix + ".conv2.weight", prefix + ".bn2", _lib.ACT_RELU, res=res)
            return Ho

        if fuse_stem:
            p10 = "head.layer1.0"
            s1, b1 = self._bn(p10 + ".bn1")
            ss, bs = self._bn(p10 + ".shortcut.1")
            H = F // 2
            prog.op(_lib.OP_STEM_BLOCK, in_buf=0, out_buf=bb, out_ld=C0, res_buf=bc, res_ld=C0, H=F, W=T, Ho=H, Wo=T, Cin=1, Cout=C0,
                    w=self._raw("head.conv1.weight"), epi_scale=s, epi_shift=b, act=_lib.ACT_RELU,
                    aux=[self._w2d(p10 + ".conv1.weight"), s1, b1, self._w2d(p10 + ".shortcut.0.weight")], iaux=[ss, bs])
            conv3x3(bb, bd, H, 1, p10 + ".conv2.weight", p10 + ".bn2", _lib.ACT_RELU, res=bc)
        else:
            H = res_block("head.layer1.0", big, F, 2, bb, bc, bd)        # -> bd  [F/2]
        H = res_block("head.layer1.1", bd, H, 1, bb, -1, bc)         # -> bc
        H = res_block("head.layer2.0", bc, H, 2, bb, bd, big)        # -> big [F/4]
        H = res_block("head.layer2.1", big, H, 1, bb, -1, bd)        # -> bd
        fo = prog.buf("fcm_out", (F // 8) * T * C0, AD)               # crosses into phase 1: exact size
        H = conv3x3(bd, fo, H, 2, "head.conv2.weight", "head.bn2", _lib.ACT_RELU)   # -> [F/8, T, 32]
        assert H == F // 8
        bb = fo
        prog.phase = 1        # D-TDNN part: small per-segment activations, run over the coarse sub-batch

        # xvector
        T2 = conv_out(T, 5, 2, 2)
        nwin = (T2 + SEG_LEN - 1) // SEG_LEN
        G, BNC = mod.growth_rate, mod.bn_channels
        widths = []
        ch = mod.init_channels
        for (n_layers, _, _) in mod.block_cfg:
            widths.append(ch + n_layers * G)
            ch = (ch + n_layers * G) // 2
        xbufs = [prog.buf("block%d" % (i + 1), T2 * w, AD) for i, w in enumerate(widths)]
        hbuf = prog.buf("bottleneck", T2 * BNC, AD)
        gbuf = prog.buf("gate", nwin * G, _lib.DT_F32)
        fin = prog.buf("final", T2 * mod.final_channels, _lib.DT_F32)   # fp32 into the std pooling: bf16
        #   quantisation of near-constant channels would dominate their (tiny) standard deviation
        stats = prog.buf("stats", 2 * mod.final_channels, _lib.DT_F32)

        # tdnn: Conv1d(320->128,k5,s2,p2) over channels c*H+f == conv over the [H,T,32] map with KH=H
        def tdnn_w():
            w = self.sd["xvector.tdnn.linear.weight"]               # [128, 32*H, 5]
            return w.reshape(w.shape[0], C0, H, 5).permute(0, 2, 3, 1).contiguous()
        sc, sh = self._bn("xvector.tdnn.nonlinear.batchnorm")
        prog.op(_lib.OP_CONV, in_buf=bb, in_ld=C0, out_buf=xbufs[0], out_ld=widths[0], H=H, W=T, Cin=C0, Ho=1,
                Wo=T2, Cout=mod.init_channels, KH=H, KW=5, sh=1, sw=2, ph=0, pw=2,
                w=self._p(("w", "tdnn"), tdnn_w), epi_scale=sc, epi_shift=sh, act=_lib.ACT_RELU)

        ch = mod.init_channels
        for bi, (n_layers, k, dil) in enumerate(mod.block_cfg):
            xb, ld = xbufs[bi], widths[bi]
            for j in range(n_layers):
                p = "xvector.block%d.tdnnd%d" % (bi + 1, j + 1)
                cin = ch + j * G
                ps, pb = self._bn(p + ".nonlinear1.batchnorm")
                es, eb = self._bn(p + ".nonlinear2.batchnorm")
                # L2 hints: the concat buffer streams through (evict_first), the 128-channel bottleneck output is read back
                # by the CAM layer right away and overwritten by the next layer's: it should never leave the L2
                prog.op(_lib.OP_CONV, in_buf=xb, in_ld=ld, out_buf=hbuf, out_ld=BNC, H=1, W=T2, Cin=cin, Ho=1, Wo=T2,
                        Cout=BNC, w=self._w1d(p + ".linear1.weight"), pro_scale=ps, pro_shift=pb, pro_relu=1,
                        epi_scale=es, epi_shift=eb, act=_lib.ACT_RELU, reserved=L2_HINTS)
                c = p + ".cam_layer"
                # whole CAMLayer as one op: context gate + dilated local conv + gating, written in place
                # into the block's concat buffer (fused kernel in bf16 mode, gate + gated conv otherwise)
                prog.op(_lib.OP_CAM_LOCAL, in_buf=hbuf, in_ld=BNC, out_buf=xb, out_ld=ld, out_choff=cin, H=1, W=T2,
                        Cin=BNC, Ho=1, Wo=T2, Cout=G, KH=1, KW=k, pw=dil * (k - 1) // 2, dw=dil,
                        w=self._w1d(c + ".linear_local.weight"), gate_buf=gbuf, gate_win=SEG_LEN,
                        aux=[self._raw(c + ".linear1.weight"), self._raw(c + ".linear1.bias"),
                             self._raw(c + ".linear2.weight"), self._raw(c + ".linear2.bias")],
                        iaux=[BNC // 2, SEG_LEN,
                              self._p(("w1t", c), lambda c=c: self.sd[c + ".linear1.weight"][:, :, 0].t().contiguous()),
                              self._p(("w2t", c), lambda c=c: self.sd[c + ".linear2.weight"][:, :, 0].t().contiguous())])
            ch = ch + n_layers * G
            p = "xvector.transit%d" % (bi + 1)
            ps, pb = self._bn(p + ".nonlinear.batchnorm")
            last = bi == len(mod.block_cfg) - 1
            kw = dict(in_buf=xb, in_ld=ld, H=1, W=T2, Cin=ch, Ho=1, Wo=T2, Cout=ch // 2,
                      w=self._w1d(p + ".linear.weight"), pro_scale=ps, pro_shift=pb, pro_relu=1)
            if last:
                es, eb = self._bn("xvector.out_nonlinear.batchnorm")
                kw.update(out_buf=fin, out_ld=ch // 2, epi_scale=es, epi_shift=eb, act=_lib.ACT_RELU)
            else:
                kw.update(out_buf=xbufs[bi + 1], out_ld=widths[bi + 1])
            prog.op(_lib.OP_CONV, **kw)
            ch //= 2
        prog.op(_lib.OP_STATS_POOL, in_buf=fin, in_ld=ch, out_buf=stats, H=1, W=T2, Cin=ch, iaux=[1], faux=[0.0])
        es, eb = self._bn("xvector.dense.nonlinear.batchnorm", affine=False)
        prog.op(_lib.OP_CONV, in_buf=stats, in_ld=2 * ch, out_buf=1, out_ld=E, H=1, W=1, Cin=2 * ch, Ho=1, Wo=1,
                Cout=E, w=self._w1d("xvector.dense.linear.weight"), epi_scale=es, epi_shift=eb)
        self.model.set_program(T, prog)


CAMPPlus.engine_cls = _Engine
