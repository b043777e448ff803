"""Drop-in for ``speakerlab.models.eres2net.ERes2Net.ERes2Net`` (ERes2Net.py:154-231; base 6.6 M and, with
``m_channels=64``, large) and ``speakerlab.models.eres2net.ERes2Net_huge.ERes2Net`` (``ERes2Net_huge`` below: the same
graph with expansion 4, baseWidth 24, scale 3, m_channels 64).

Same constructor signature (``block`` / ``block_fuse`` are accepted and ignored: the two block types are fixed by the
architecture; the block hyper-parameters the reference hard-codes in its block classes are keyword arguments here)
and the same ``state_dict`` layout, so reference checkpoints load unchanged.  ``forward`` compiles the eval-mode
network into the fused-op list of ``libb200spk`` with the ERes2NetV2 machinery - the residual blocks are the same
ops - plus the bottom-up fusion chain of v1:

  layer1_downsample / 2 / 3 (3x3 stride 2, no BN)   ERes2Net.py:179-181, 213-220   CONV
  fuse_mode12 / 123 / 1234 (AFF)                     ERes2Net.py:184-186, 214-221   CONV x3 + AFF_BLEND
  TSTP + seg_1                                       ERes2Net.py:222-224            STATS_POOL + CONV 1x1

Each downsample runs as soon as its input exists (right after the layer that produces it), so no layer output has to
outlive the next layer.  Inference only.
"""
from torch import nn

from . import _lib
from .eres2netv2 import _Engine as _V2Engine, _aff, _block, _pad16
from .program import EngineModule, conv_out


class ERes2Net(EngineModule):
    def __init__(self, block=None, block_fuse=None, num_blocks=[3, 4, 6, 3], m_channels=32, feat_dim=80, embedding_size=192,
                 pooling_func='TSTP', two_emb_layer=False, baseWidth=32, scale=2, expansion=2, precision="fp32", chunk=None):
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
        X = expansion
        self.layer1_downsample = nn.Conv2d(m_channels * X, m_channels * 2 * X, 3, stride=2, padding=1, bias=False)
        self.layer2_downsample = nn.Conv2d(m_channels * 2 * X, m_channels * 4 * X, 3, padding=1, stride=2, bias=False)
        self.layer3_downsample = nn.Conv2d(m_channels * 4 * X, m_channels * 8 * X, 3, padding=1, stride=2, bias=False)
        self.fuse_mode12 = _aff(m_channels * 2 * X, 4)
        self.fuse_mode123 = _aff(m_channels * 4 * X, 4)
        self.fuse_mode1234 = _aff(m_channels * 8 * X, 4)
        self.seg_1 = nn.Linear(self.stats_dim * expansion * 2, embedding_size)
        self.eval()

    def forward(self, x):
        """x [B, T, feat_dim] -> [B, embedding_size] (float32, same device); the caller's tensor is not modified."""
        return self._run(x, self.feat_dim, self.embedding_size)


def ERes2Net_huge(num_blocks=[3, 4, 6, 3], m_channels=64, feat_dim=80, embedding_size=192, pooling_func='TSTP',
                  two_emb_layer=False, precision="fp32", chunk=None, **ignored):
    """``speakerlab.models.eres2net.ERes2Net_huge.ERes2Net``: expansion 4, baseWidth 24, scale 3 (ERes2Net_huge.py:32-34)."""
    return ERes2Net(num_blocks=num_blocks, m_channels=m_channels, feat_dim=feat_dim, embedding_size=embedding_size,
                    pooling_func=pooling_func, two_emb_layer=two_emb_layer, baseWidth=24, scale=3, expansion=4,
                    precision=precision, chunk=chunk)


class _Engine(_V2Engine):
    def keep_layers(self):
        return set()

    def _fuse(self, env, name, x, y, C, H, W, out_dtype=None):
        """AFF(x, y) on [H*W, C] maps -> its own buffer."""
        ip = _pad16(C // 4)
        dt = {} if out_dtype is None else {"dt": out_dtype}
        out = env.buf(name, H * W * C, **dt)
        env.aff_ops(name.replace("fuse", "fuse_mode"), x, C, 0, y, C, 0, C, C, H, W, env.buf(name + "_raw", H * W * ip),
                    env.buf(name + "_t", H * W * ip), env.buf(name + "_z", H * W * C), out, C, 0)
        return out

    def _down(self, env, key, src, cin, H, W, cout):
        Ho, Wo = conv_out(H, 3, 2, 1), conv_out(W, 3, 2, 1)
        dst = env.buf(key, Ho * Wo * cout)
        env.prog.op(_lib.OP_CONV, in_buf=src, in_ld=cin, out_buf=dst, out_ld=cout, H=H, W=W, Cin=cin, Ho=Ho, Wo=Wo, Cout=cout,
                    KH=3, KW=3, sh=2, sw=2, ph=1, pw=1, w=self._w2d(key + ".weight"))
        return dst, Ho, Wo

    def after_layer(self, li, env):
        out, c, H, W = env.layer_out[li]
        if li == 1:
            self._ds = self._down(env, "layer1_downsample", out, c, H, W, 2 * c)
        elif li in (2, 3):
            ds, Hd, Wd = self._ds
            assert (Hd, Wd) == (H, W)
            fused = self._fuse(env, "fuse12" if li == 2 else "fuse123", out, ds, c, H, W)
            self._ds = self._down(env, "layer%d_downsample" % li, fused, c, H, W, 2 * c)

    def tail(self, env):
        out4, c4, H4, W4 = env.layer_out[4]
        ds, Hd, Wd = self._ds
        assert (Hd, Wd) == (H4, W4)
        fused = self._fuse(env, "fuse1234", out4, ds, c4, H4, W4, _lib.DT_F32)      # fp32 into the std pooling
        self.pool_and_embed(env, fused, c4, H4, W4)


ERes2Net.engine_cls = _Engine
