"""Plain-PyTorch fp32 CPU restatement of the ERes2Net (v1: base / large / huge) forward, driven by a state_dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional, eval-mode BatchNorm, written against the reference's
state_dict key names.  Follows:

  ERes2Net.forward                     speakerlab/models/eres2net/ERes2Net.py:207-231  (ERes2Net_huge.py: same graph)
  BasicBlockERes2Net.forward           ERes2Net.py:60-88        (the blocks are the V2 blocks without the V2 width rule)
  BasicBlockERes2Net_diff_AFF.forward  ERes2Net.py:126-152
  AFF.forward                          speakerlab/models/eres2net/fusion.py:22-28
  TSTP.forward                         speakerlab/models/eres2net/pooling_layers.py:47-55

Pinned by tests/golden/eres2net.npz (minted from the imported reference by oracle/gen_golden.py eres2net).
"""
import torch
import torch.nn.functional as F

from .eres2netv2_oracle import _t, aff, block, _bn


def forward(sd, feats, num_blocks=(3, 4, 6, 3), scale=2, taps=None):
    """feats [B,T,80] float32 -> embeddings [B,E].  ``scale`` = 2 (base / large) or 3 (huge)."""
    with torch.no_grad():
        x = torch.as_tensor(feats, dtype=torch.float32).permute(0, 2, 1).unsqueeze(1)
        out = F.relu(_bn(sd, "bn1", F.conv2d(x, _t(sd, "conv1.weight"), padding=1)))
        outs = []
        for li, (nb, stride, fuse) in enumerate(zip(num_blocks, (1, 2, 2, 2), (False, False, True, True)), start=1):
            for bi in range(nb):
                out = block(sd, "layer%d.%d" % (li, bi), out, stride if bi == 0 else 1, scale, fuse)
            outs.append(out)
            if taps is not None:
                taps["layer%d" % li] = out
        down = lambda name, t: F.conv2d(t, _t(sd, name + ".weight"), stride=2, padding=1)      # noqa: E731
        f12 = aff(sd, "fuse_mode12", outs[1], down("layer1_downsample", outs[0]))
        f123 = aff(sd, "fuse_mode123", outs[2], down("layer2_downsample", f12))
        f1234 = aff(sd, "fuse_mode1234", outs[3], down("layer3_downsample", f123))
        if taps is not None:
            taps.update(fuse12=f12, fuse123=f123, fuse1234=f1234)
        mean = f1234.mean(dim=-1).flatten(1)
        std = torch.sqrt(torch.var(f1234, dim=-1) + 1e-8).flatten(1)
        return F.linear(torch.cat((mean, std), 1), _t(sd, "seg_1.weight"), _t(sd, "seg_1.bias"))
