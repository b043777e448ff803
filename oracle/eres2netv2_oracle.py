"""Plain-PyTorch fp32 CPU restatement of the ERes2NetV2 forward, driven by a state_dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional, eval-mode BatchNorm, written against
the reference's state_dict key names.  Follows:

  ERes2NetV2.forward             speakerlab/models/eres2net/ERes2NetV2.py:235-254
  BasicBlockERes2NetV2.forward   speakerlab/models/eres2net/ERes2NetV2.py:65-91
  BasicBlockERes2NetV2AFF.forward speakerlab/models/eres2net/ERes2NetV2.py:132-159
  AFF.forward                    speakerlab/models/eres2net/fusion.py:22-28
  ReLU = Hardtanh(0, 20)         speakerlab/models/eres2net/ERes2NetV2.py:20-28
  TSTP.forward                   speakerlab/models/eres2net/pooling_layers.py:47-55

Pinned by tests/golden/eres2netv2.npz (minted from the imported reference).
"""
import math

import torch
import torch.nn.functional as F


def _t(sd, key):
    v = sd[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(v)


def _bn(sd, p, x):
    return F.batch_norm(x, _t(sd, p + ".running_mean"), _t(sd, p + ".running_var"), _t(sd, p + ".weight"),
                        _t(sd, p + ".bias"), training=False, eps=1e-5)


def _clamp(x):
    return x.clamp(0.0, 20.0)


def aff(sd, p, x, y):
    xa = torch.cat((x, y), dim=1)
    h = F.conv2d(xa, _t(sd, p + ".local_att.0.weight"), _t(sd, p + ".local_att.0.bias"))
    h = F.silu(_bn(sd, p + ".local_att.1", h))
    h = _bn(sd, p + ".local_att.4", F.conv2d(h, _t(sd, p + ".local_att.3.weight"), _t(sd, p + ".local_att.3.bias")))
    g = 1.0 + torch.tanh(h)
    return x * g + y * (2.0 - g)


def block(sd, p, x, stride, scale, fuse):
    width = _t(sd, p + ".convs.0.weight").shape[0]
    out = _clamp(_bn(sd, p + ".bn1", F.conv2d(x, _t(sd, p + ".conv1.weight"), stride=stride)))
    spx = torch.split(out, width, 1)
    outs = []
    sp = None
    for i in range(scale):
        if i == 0:
            sp = spx[0]
        elif fuse:
            sp = aff(sd, p + ".fuse_models.%d" % (i - 1), sp, spx[i])
        else:
            sp = sp + spx[i]
        sp = _clamp(_bn(sd, p + ".bns.%d" % i, F.conv2d(sp, _t(sd, p + ".convs.%d.weight" % i), padding=1)))
        outs.append(sp)
    out = _bn(sd, p + ".bn3", F.conv2d(torch.cat(outs, 1), _t(sd, p + ".conv3.weight")))
    if (p + ".shortcut.0.weight") in sd:
        res = _bn(sd, p + ".shortcut.1", F.conv2d(x, _t(sd, p + ".shortcut.0.weight"), stride=stride))
    else:
        res = x
    return _clamp(out + res)


def forward(sd, feats, num_blocks=(3, 4, 6, 3), scale=2, taps=None):
    """feats [B,T,80] float32 -> embeddings [B,E]."""
    with torch.no_grad():
        x = torch.as_tensor(feats, dtype=torch.float32).permute(0, 2, 1).unsqueeze(1)
        out = F.relu(_bn(sd, "bn1", F.conv2d(x, _t(sd, "conv1.weight"), padding=1)))
        feats_l = []
        for li, (nb, stride, fuse) in enumerate(zip(num_blocks, (1, 2, 2, 2), (False, False, True, True)), start=1):
            for bi in range(nb):
                out = block(sd, "layer%d.%d" % (li, bi), out, stride if bi == 0 else 1, scale, fuse)
            feats_l.append(out)
            if taps is not None:
                taps["layer%d" % li] = out
        out3_ds = F.conv2d(feats_l[2], _t(sd, "layer3_ds.weight"), stride=2, padding=1)
        fused = aff(sd, "fuse34", feats_l[3], out3_ds)
        if taps is not None:
            taps["fuse34"] = fused
        mean = fused.mean(dim=-1).flatten(1)
        std = torch.sqrt(torch.var(fused, dim=-1) + 1e-8).flatten(1)
        stats = torch.cat((mean, std), 1)
        return F.linear(stats, _t(sd, "seg_1.weight"), _t(sd, "seg_1.bias"))
