"""Drop-in for ``speakerlab.models.eres2net.ERes2NetV2.ERes2NetV2`` (ERes2NetV2.py:161-254).

Same constructor signature (``block`` / ``block_fuse`` are accepted and ignored: the two block
types are fixed by the architecture) and the same ``state_dict`` layout (557 keys for the default
``baseWidth=26, scale=2, expansion=2``; 1001 for ``24, 4, 4``).  ``forward`` compiles the eval-mode
network into the fused-op list of ``libb200spk``.  Inference only.

Channel padding: the Res2Net group width ``floor(planes*baseWidth/64)`` (26/52/104/208, or
24/48/96/192) is padded to a multiple of 16 inside the packed weights and folded-BN vectors
(zero rows/columns, zero scale/shift), so padded channels carry exact zeros through
clamp(0,20), the 3x3 convs, AFF and the 1x1 projections.

Op mapping (reference file:line -> fused op):
  conv1 + bn1 + relu                       :236-238      STEM
  block conv1 (1x1, stride s) + bn1 + clamp :66-68       CONV
  hierarchical sum sp + spx[i]              :74-75       AFF_BLEND without z (plain add)
  AFF(sp, spx[i]) / fuse34                  fusion.py:22-28   CONV(x half, BN scale folded) -> CONV(y half, +res, SiLU) -> CONV(+BN) -> AFF_BLEND
  convs[i] 3x3 + bns[i] + clamp, cat        :76-82       CONV writing its channel slice of the concat buffer
  conv3 + bn3 + shortcut + clamp            :84-90       CONV(shortcut+BN) / CONV(+BN, +res, clamp)
  layer3_ds 3x3 stride 2                    :222-223,244 CONV
  TSTP                                      pooling_layers.py:47-55   STATS_POOL (per frequency row, unbiased, eps 1e-8)
  seg_1 Linear                              :218-219,247 CONV 1x1 with the weight columns permuted from (c,f) to (f,c) order
"""
import math

import torch
from torch import nn

from . import _lib
from .program import EngineBase, EngineModule, Program, conv_out, fold_bn


def _pad16(c):
    return (c + 15) // 16 * 16


class _Holder(nn.Module):
    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.add_module(k, v)


def _aff(channels, r=4):
    inter = int(channels // r)
    return _Holder(local_att=nn.Sequential(nn.Conv2d(channels * 2, inter, 1), nn.BatchNorm2d(inter), nn.SiLU(inplace=True),
                                           nn.Conv2d(inter, channels, 1), nn.BatchNorm2d(channels)))


def _block(in_planes, planes, stride, base_width, scale, expansion, fuse):
    width = int(math.floor(planes * (base_width / 64.0)))
    kids = dict(conv1=nn.Conv2d(in_planes, width * scale, 1, stride=stride, bias=False), bn1=nn.BatchNorm2d(width * scale),
                convs=nn.ModuleList([nn.Conv2d(width, width, 3, padding=1, bias=False) for _ in range(scale)]),
                bns=nn.ModuleList([nn.BatchNorm2d(width) for _ in range(scale)]))
    if fuse:
        kids["fuse_models"] = nn.ModuleList([_aff(width, 4) for _ in range(scale - 1)])
    kids["conv3"] = nn.Conv2d(width * scale, planes * expansion, 1, bias=False)
    kids["bn3"] = nn.BatchNorm2d(planes * expansion)
    sc = nn.Sequential()
    if stride != 1 or in_planes != expansion * planes:
        sc = nn.Sequential(nn.Conv2d(in_planes, expansion * planes, 1, stride=stride, bias=False),
                           nn.BatchNorm2d(expansion * planes))
    kids["shortcut"] = sc
    blk = _Holder(**kids)
    blk.cfg = (in_planes, planes, stride, width, fuse)
    return blk


class ERes2NetV2(EngineModule):
    def __init__(self, block=None, block_fuse=None, num_blocks=[3, 4, 6, 3], m_channels=64, feat_dim=80,
                 embedding_size=192, baseWidth=26, scale=2, expansion=2, pooling_func='TSTP', two_emb_layer=False,
                 precision="fp32", chunk=None):
        super().__init__()
        assert pooling_func == 'TSTP' and not two_emb_layer, "only the shipped TSTP / single-embedding head is implemented"
        assert feat_dim % 8 == 0
        self._init_engine(precision, chunk)
        self.feat_dim, self.embedding_size = feat_dim, embedding_size
        self.m_channels, self.baseWidth, self.scale, self.expansion = m_channels, baseWidth, scale, expansion
        self.num_blocks = list(num_blocks)
        self.stats_dim = int(feat_dim / 8) * m_channels * 8
        self.conv1 = nn.Conv2d(1, m_channels, 3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(m_channels)
        in_planes = m_channels
        for li, (mult, nb, stride, fuse) in enumerate(zip((1, 2, 4, 8), self.num_blocks, (1, 2, 2, 2),
                                                          (False, False, True, True)), start=1):
            blocks = []
            for s in [stride] + [1] * (nb - 1):
                blocks.append(_block(in_planes, m_channels * mult, s, baseWidth, scale, expansion, fuse))
                in_planes = m_channels * mult * expansion
            self.add_module("layer%d" % li, nn.Sequential(*blocks))
        self.layer3_ds = nn.Conv2d(m_channels * 4 * expansion, m_channels * 8 * expansion, 3, padding=1, stride=2, bias=False)
        self.fuse34 = _aff(m_channels * 8 * expansion, 4)
        self.seg_1 = nn.Linear(self.stats_dim * expansion * 2, embedding_size)
        self.eval()

    def forward(self, x):
        """x [B, T, feat_dim] -> [B, embedding_size] (float32, same device).  Unlike the reference
        (ERes2NetV2.py:237 ``unsqueeze_``) the caller's tensor is not modified."""
        return self._run(x, self.feat_dim, self.embedding_size)


class _Engine(EngineBase):
    def default_chunks(self, T):
        # ~ (6 MB default / 20 MB w24s4ep4) of bf16 activations per 3 s segment: modest sub-batches
        per_seg = 80 * T * 64 * self.m.expansion * (2 if self.model.precision == _lib.PREC_BF16 else 4) * 6
        c = max(4, min(256, int(6e9 // per_seg)))
        return c, c

    # ---- padded parameter builders
    def _vec(self, key, groups, w, wp, fill=0.0):
        """[groups*w] vector -> [groups*wp], each group zero-padded (fill for the padding)."""
        def build():
            v = self.sd[key] if isinstance(key, str) else key()
            out = torch.full((groups, wp), fill, dtype=torch.float32)
            out[:, :w] = v.reshape(groups, w)
            return out.reshape(-1)
        return build

    def _bn_padded(self, prefix, groups, w, wp):
        s = self._p(("bnp_s", prefix), self._vec(lambda: fold_bn(self.sd, prefix)[0], groups, w, wp))
        b = self._p(("bnp_b", prefix), self._vec(lambda: fold_bn(self.sd, prefix)[1], groups, w, wp))
        return s, b

    def _w_rows_cols(self, key, rg, rw, rwp, cg, cw, cwp):
        """conv weight [rg*rw, cg*cw, KH, KW] -> packed [rg*rwp, KH, KW, cg*cwp] with zero padding."""
        def build():
            w = self.sd[key]
            kh, kw = w.shape[2], w.shape[3]
            out = torch.zeros((rg, rwp, kh, kw, cg, cwp), dtype=torch.float32)
            out[:, :rw, :, :, :, :cw] = w.reshape(rg, rw, cg, cw, kh, kw).permute(0, 1, 4, 5, 2, 3)
            return out.reshape(rg * rwp, kh, kw, cg * cwp)
        return self._p(("wpad", key, rwp, cwp), build)

    def compile(self, T):
        mod, AD = self.m, self.model.act_dtype
        F, E, S, X = mod.feat_dim, mod.embedding_size, mod.scale, mod.expansion
        C0 = mod.m_channels
        prog = Program(T * F, E)
        bufs = {}

        def buf(name, elems, dt=AD):
            # buffers are reused by name across blocks; sized for the largest request
            if name in bufs and bufs[name][1] >= elems:
                return bufs[name][0]
            assert name not in bufs, "buffer %s requested with a larger size later (%d)" % (name, elems)
            bufs[name] = (prog.buf(name, elems, dt), elems)
            return bufs[name][0]

        H, W = F, T
        max_hw = H * W
        stem = buf("stem", max_hw * C0)
        s, b = self._bn("bn1")
        prog.op(_lib.OP_STEM, in_buf=0, out_buf=stem, out_ld=C0, H=F, W=T, Cout=C0, w=self._raw("conv1.weight"),
                epi_scale=s, epi_shift=b, act=_lib.ACT_RELU)
        # two ping-pong buffers for block outputs, sized for layer1's output
        c_l1 = C0 * X
        ping = [buf("blk_a", max_hw * max(c_l1, C0)), buf("blk_b", max_hw * max(c_l1, C0))]
        cur, cur_c, which = stem, C0, 0
        layer_out = {}
        # scratch sized for the widest use (layer1 geometry has the most pixels)
        widths = [int(math.floor(C0 * m * (mod.baseWidth / 64.0))) for m in (1, 2, 4, 8)]
        wp_all = [_pad16(w) for w in widths]
        hw_l = []
        h_, w_ = F, T
        for st in (1, 2, 2, 2):
            h_, w_ = conv_out(h_, 1, st, 0), conv_out(w_, 1, st, 0)
            hw_l.append(h_ * w_)
        spx_b = buf("spx", max(hw * S * wp for hw, wp in zip(hw_l, wp_all)))
        cat_b = buf("cat", max(hw * S * wp for hw, wp in zip(hw_l, wp_all)))
        sum_b = buf("sum", max(hw * wp for hw, wp in zip(hw_l, wp_all)))
        sc_b = buf("shortcut", max(hw * C0 * m * X for hw, m in zip(hw_l, (1, 2, 4, 8))))
        in_blk = [(hw, w, wp) for hw, w, wp in zip(hw_l[2:], widths[2:], wp_all[2:])]
        aff_raw = buf("aff_raw", max(hw * _pad16(max(1, w // 4)) for hw, w, _ in in_blk))
        aff_t = buf("aff_t", max(hw * _pad16(max(1, w // 4)) for hw, w, _ in in_blk))
        aff_z = buf("aff_z", max(hw * wp for hw, _, wp in in_blk))
        keep = self.keep_layers()      # layers whose output must survive later layers: own buffer
        env = type("Env", (), {})()

        def aff_ops(prefix, x_buf, x_ld, x_off, y_buf, y_ld, y_off, C, Cp, hw_h, hw_w, t_raw, t_buf, z_buf, out_buf, out_ld, out_off):
            """AFF (fusion.py:22-28) on x,y [hw, Cp] (C real channels) -> out."""
            inter = int(C // 4)
            ip = _pad16(inter)
            la = prefix + ".local_att"
            s1, h1 = fold_bn(self.sd, la + ".1")
            w0, b0 = self.sd[la + ".0.weight"], self.sd[la + ".0.bias"]

            def half(lo):
                def build():
                    out = torch.zeros((ip, 1, 1, Cp))
                    out[:inter, 0, 0, :C] = w0[:, lo:lo + C, 0, 0]
                    return out
                return build
            wx = self._p(("affx", prefix), half(0))
            wy = self._p(("affy", prefix), half(C))
            s1p = self._p(("aff_s1", prefix), lambda: torch.cat([s1, torch.zeros(ip - inter)]))
            zero_ip = self._p(("zeros", ip), lambda: torch.zeros(ip))
            h1p = self._p(("aff_h1", prefix), lambda: torch.cat([s1 * b0 + h1, torch.zeros(ip - inter)]))
            # t_raw = s1 * (Wx x); t = SiLU(s1 * (Wy y + b0) + h1 + t_raw): the 1x1 over cat(x, y) as two GEMMs
            prog.op(_lib.OP_CONV, in_buf=x_buf, in_ld=x_ld, in_choff=x_off, out_buf=t_raw, out_ld=ip, H=hw_h, W=hw_w,
                    Cin=Cp, Ho=hw_h, Wo=hw_w, Cout=ip, w=wx, epi_scale=s1p, epi_shift=zero_ip)
            prog.op(_lib.OP_CONV, in_buf=y_buf, in_ld=y_ld, in_choff=y_off, out_buf=t_buf, out_ld=ip, H=hw_h, W=hw_w,
                    Cin=Cp, Ho=hw_h, Wo=hw_w, Cout=ip, w=wy, epi_scale=s1p, epi_shift=h1p, res_buf=t_raw, res_ld=ip,
                    act=_lib.ACT_SILU)
            s2, h2 = fold_bn(self.sd, la + ".4")
            w3, b3 = self.sd[la + ".3.weight"], self.sd[la + ".3.bias"]

            def w3p():
                out = torch.zeros((Cp, 1, 1, ip))
                out[:C, 0, 0, :inter] = w3[:, :, 0, 0]
                return out
            s2p = self._p(("aff_s2", prefix), lambda: torch.cat([s2, torch.zeros(Cp - C)]))
            h2p = self._p(("aff_h2", prefix), lambda: torch.cat([s2 * b3 + h2, torch.zeros(Cp - C)]))
            prog.op(_lib.OP_CONV, in_buf=t_buf, in_ld=ip, out_buf=z_buf, out_ld=Cp, H=hw_h, W=hw_w, Cin=ip, Ho=hw_h,
                    Wo=hw_w, Cout=Cp, w=self._p(("aff_w3", prefix), w3p), epi_scale=s2p, epi_shift=h2p)
            prog.op(_lib.OP_AFF_BLEND, in_buf=x_buf, in_ld=x_ld, in_choff=x_off, res_buf=y_buf, res_ld=y_ld, res_choff=y_off,
                    gate_buf=z_buf, iaux=[Cp, 0], out_buf=out_buf, out_ld=out_ld, out_choff=out_off, H=hw_h, W=hw_w, Cin=Cp)

        for li in range(1, 5):
            layer = getattr(mod, "layer%d" % li)
            for bi, blk in enumerate(layer):
                in_planes, planes, stride, width, fuse = blk.cfg
                p = "layer%d.%d" % (li, bi)
                wp = _pad16(width)
                Ho, Wo = conv_out(H, 1, stride, 0), conv_out(W, 1, stride, 0)
                hw = Ho * Wo
                cout = planes * X
                # conv1 (1x1, stride on both axes) + bn1 + clamp -> spx [hw, S*wp]
                s1, b1 = self._bn_padded(p + ".bn1", S, width, wp)
                prog.op(_lib.OP_CONV, in_buf=cur, in_ld=cur_c, out_buf=spx_b, out_ld=S * wp, H=H, W=W, Cin=cur_c, Ho=Ho, Wo=Wo,
                        Cout=S * wp, sh=stride, sw=stride, w=self._w_rows_cols(p + ".conv1.weight", S, width, wp, 1, cur_c, cur_c),
                        epi_scale=s1, epi_shift=b1, act=_lib.ACT_CLAMP20)
                for i in range(S):
                    if i == 0:
                        src, src_ld, src_off = spx_b, S * wp, 0
                    elif fuse:
                        aff_ops("%s.fuse_models.%d" % (p, i - 1), cat_b, S * wp, (i - 1) * wp, spx_b, S * wp, i * wp, width, wp,
                                Ho, Wo, aff_raw, aff_t, aff_z, sum_b, wp, 0)
                        src, src_ld, src_off = sum_b, wp, 0
                    else:
                        prog.op(_lib.OP_AFF_BLEND, in_buf=cat_b, in_ld=S * wp, in_choff=(i - 1) * wp, res_buf=spx_b, res_ld=S * wp,
                                res_choff=i * wp, out_buf=sum_b, out_ld=wp, H=Ho, W=Wo, Cin=wp)
                        src, src_ld, src_off = sum_b, wp, 0
                    si, bi_ = self._bn_padded("%s.bns.%d" % (p, i), 1, width, wp)
                    prog.op(_lib.OP_CONV, in_buf=src, in_ld=src_ld, in_choff=src_off, out_buf=cat_b, out_ld=S * wp, out_choff=i * wp,
                            H=Ho, W=Wo, Cin=wp, Ho=Ho, Wo=Wo, Cout=wp, KH=3, KW=3, ph=1, pw=1,
                            w=self._w_rows_cols("%s.convs.%d.weight" % (p, i), 1, width, wp, 1, width, wp),
                            epi_scale=si, epi_shift=bi_, act=_lib.ACT_CLAMP20)
                # shortcut
                if (p + ".shortcut.0.weight") in self.sd:
                    ss, sb = self._bn(p + ".shortcut.1")
                    prog.op(_lib.OP_CONV, in_buf=cur, in_ld=cur_c, out_buf=sc_b, out_ld=cout, H=H, W=W, Cin=cur_c, Ho=Ho, Wo=Wo,
                            Cout=cout, sh=stride, sw=stride, w=self._w2d(p + ".shortcut.0.weight"), epi_scale=ss, epi_shift=sb)
                    res, res_ld = sc_b, cout
                else:
                    res, res_ld = cur, cur_c
                # conv3 + bn3 + residual + clamp
                if li in keep and bi == len(layer) - 1:
                    dst = buf("out%d" % li, hw * cout)
                else:
                    dst = ping[which]
                    which ^= 1
                    if dst == cur:                      # never write the buffer being read as the residual
                        dst = ping[which]
                        which ^= 1
                s3, b3 = self._bn(p + ".bn3")
                prog.op(_lib.OP_CONV, in_buf=cat_b, in_ld=S * wp, out_buf=dst, out_ld=cout, H=Ho, W=Wo, Cin=S * wp, Ho=Ho, Wo=Wo,
                        Cout=cout, w=self._w_rows_cols(p + ".conv3.weight", 1, cout, cout, S, width, wp),
                        epi_scale=s3, epi_shift=b3, res_buf=res, res_ld=res_ld, act=_lib.ACT_CLAMP20)
                cur, cur_c, H, W = dst, cout, Ho, Wo
            layer_out[li] = (cur, cur_c, H, W)
            env.prog, env.buf, env.aff_ops, env.layer_out, env.E = prog, buf, aff_ops, layer_out, E
            self.after_layer(li, env)
        self.tail(env)
        self.model.set_program(T, prog)

    def keep_layers(self):
        return {3}                      # layer3_ds reads the layer-3 output after layer 4 ran

    def after_layer(self, li, env):
        pass

    def tail(self, env):
        """layer3_ds -> fuse34 -> TSTP -> seg_1 (ERes2NetV2.py:244-247)."""
        prog, buf, aff_ops, layer_out, E = env.prog, env.buf, env.aff_ops, env.layer_out, env.E
        out4, c4, H4, W4 = layer_out[4]
        o3, c3, H3, W3 = layer_out[3]
        # layer3_ds: 3x3 stride 2 pad 1, no BN / activation
        ds_b = buf("out3_ds", H4 * W4 * c4)
        assert conv_out(H3, 3, 2, 1) == H4 and conv_out(W3, 3, 2, 1) == W4
        prog.op(_lib.OP_CONV, in_buf=o3, in_ld=c3, out_buf=ds_b, out_ld=c4, H=H3, W=W3, Cin=c3, Ho=H4, Wo=W4, Cout=c4, KH=3, KW=3,
                sh=2, sw=2, ph=1, pw=1, w=self._w2d("layer3_ds.weight"))
        fused = buf("fuse34", H4 * W4 * c4, _lib.DT_F32)       # fp32 into the std pooling
        ip34 = _pad16(c4 // 4)
        f_raw, f_t, f_z = buf("f34_raw", H4 * W4 * ip34), buf("f34_t", H4 * W4 * ip34), buf("f34_z", H4 * W4 * c4)
        aff_ops("fuse34", out4, c4, 0, ds_b, c4, 0, c4, c4, H4, W4, f_raw, f_t, f_z, fused, c4, 0)
        self.pool_and_embed(env, fused, c4, H4, W4)

    def pool_and_embed(self, env, fused, c4, H4, W4):
        """TSTP over time for every (frequency row, channel); seg_1 with columns permuted (c,f) -> (f,c)."""
        prog, buf, E = env.prog, env.buf, env.E
        stats = buf("stats", 2 * H4 * c4, _lib.DT_F32)
        prog.op(_lib.OP_STATS_POOL, in_buf=fused, in_ld=c4, out_buf=stats, H=H4, W=W4, Cin=c4, iaux=[1], faux=[1e-8])

        def seg_w():
            w = self.sd["seg_1.weight"]                                  # [E, 2*c4*H4], index = half*(c4*H4) + c*H4 + f
            w = w.reshape(E, 2, c4, H4).permute(0, 1, 3, 2).contiguous()  # -> half, f, c
            return w.reshape(E, 1, 1, 2 * H4 * c4)
        ones = self._p(("ones", E), lambda: torch.ones(E))
        prog.op(_lib.OP_CONV, in_buf=stats, in_ld=2 * H4 * c4, out_buf=1, out_ld=E, H=1, W=1, Cin=2 * H4 * c4, Ho=1, Wo=1,
                Cout=E, w=self._p(("w", "seg_1"), seg_w), epi_scale=ones, epi_shift=self._raw("seg_1.bias"))


ERes2NetV2.engine_cls = _Engine
