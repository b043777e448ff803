"""Plain-PyTorch fp32 CPU restatement of the CAM++ forward, driven by a state_dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional (no nn.Module), eval-mode
BatchNorm, written against the reference's state_dict key names so the same weights
feed the reference, this oracle and the CUDA path.  Follows:

  CAMPPlus.forward              speakerlab/models/campplus/DTDNN.py:111-115
  FCM.forward                   speakerlab/models/campplus/DTDNN.py:39-48
  BasicResBlock.forward         speakerlab/models/campplus/layers.py:248-253
  TDNNLayer                     speakerlab/models/campplus/layers.py:40-67
  CAMDenseTDNNBlock/Layer       speakerlab/models/campplus/layers.py:140-149,177-180
  CAMLayer.forward/seg_pooling  speakerlab/models/campplus/layers.py:93-110
  TransitLayer                  speakerlab/models/campplus/layers.py:193-196
  statistics_pooling            speakerlab/models/campplus/layers.py:26-32
  DenseLayer                    speakerlab/models/campplus/layers.py:209-215

Pinned by tests/golden/campplus_*.npz (minted from the imported reference).
"""
import math

import torch
import torch.nn.functional as F

BLOCKS = ((12, 1), (24, 2), (16, 2))   # (num_layers, dilation), kernel 3  DTDNN.py:77-78
SEG_LEN = 100                          # layers.py:100


def _t(sd, key):
    v = sd[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(v)


def _bn(sd, prefix, x, affine=True, eps=1e-5):
    w = _t(sd, prefix + ".weight") if affine else None
    b = _t(sd, prefix + ".bias") if affine else None
    return F.batch_norm(x, _t(sd, prefix + ".running_mean"), _t(sd, prefix + ".running_var"),
                        w, b, training=False, eps=eps)


def _res_block(sd, p, x, stride):
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, _t(sd, p + ".conv1.weight"), stride=(stride, 1), padding=1)))
    out = _bn(sd, p + ".bn2", F.conv2d(out, _t(sd, p + ".conv2.weight"), padding=1))
    if (p + ".shortcut.0.weight") in sd:
        sc = _bn(sd, p + ".shortcut.1", F.conv2d(x, _t(sd, p + ".shortcut.0.weight"), stride=(stride, 1)))
    else:
        sc = x
    return F.relu(out + sc)


def fcm(sd, x, taps=None):
    """x [B,F,T] -> [B, 32*(F/8), T]."""
    x = x.unsqueeze(1)
    out = F.relu(_bn(sd, "head.bn1", F.conv2d(x, _t(sd, "head.conv1.weight"), padding=1)))
    if taps is not None:
        taps["head.stem"] = out
    for layer in ("layer1", "layer2"):
        for i, stride in enumerate((2, 1)):
            out = _res_block(sd, "head.%s.%d" % (layer, i), out, stride)
            if taps is not None:
                taps["head.%s.%d" % (layer, i)] = out
    out = F.relu(_bn(sd, "head.bn2", F.conv2d(out, _t(sd, "head.conv2.weight"), stride=(2, 1), padding=1)))
    b, c, f, t = out.shape
    return out.reshape(b, c * f, t)


def seg_pooling(x):
    """avg_pool1d(k=100, stride=100, ceil_mode=True), expanded back to T (layers.py:100-110).
    The partial last window is averaged over its valid length."""
    t = x.shape[-1]
    n_win = math.ceil(t / SEG_LEN)
    cols = []
    for w in range(n_win):
        a, b = w * SEG_LEN, min((w + 1) * SEG_LEN, t)
        cols.append(x[..., a:b].mean(-1, keepdim=True).expand(*x.shape[:-1], b - a))
    return torch.cat(cols, dim=-1)


def _cam_layer(sd, p, x, dilation):
    y = F.conv1d(x, _t(sd, p + ".linear_local.weight"), padding=dilation, dilation=dilation)
    ctx = x.mean(-1, keepdim=True) + seg_pooling(x)
    ctx = F.relu(F.conv1d(ctx, _t(sd, p + ".linear1.weight"), _t(sd, p + ".linear1.bias")))
    m = torch.sigmoid(F.conv1d(ctx, _t(sd, p + ".linear2.weight"), _t(sd, p + ".linear2.bias")))
    return y * m


def _dense_tdnn_layer(sd, p, x, dilation):
    h = F.conv1d(F.relu(_bn(sd, p + ".nonlinear1.batchnorm", x)), _t(sd, p + ".linear1.weight"))
    h = F.relu(_bn(sd, p + ".nonlinear2.batchnorm", h))
    return _cam_layer(sd, p + ".cam_layer", h, dilation)


def forward(sd, feats, taps=None):
    """feats [B,T,80] float32 -> embeddings [B,E].  ``taps`` (dict) collects intermediate
    activations in the reference's NCHW/NCT layout for per-layer checks."""
    with torch.no_grad():
        x = torch.as_tensor(feats, dtype=torch.float32).permute(0, 2, 1)
        x = fcm(sd, x, taps)
        if taps is not None:
            taps["head"] = x
        x = F.conv1d(x, _t(sd, "xvector.tdnn.linear.weight"), stride=2, padding=2)
        x = F.relu(_bn(sd, "xvector.tdnn.nonlinear.batchnorm", x))
        if taps is not None:
            taps["xvector.tdnn"] = x
        for bi, (n_layers, dil) in enumerate(BLOCKS, start=1):
            for li in range(1, n_layers + 1):
                p = "xvector.block%d.tdnnd%d" % (bi, li)
                x = torch.cat([x, _dense_tdnn_layer(sd, p, x, dil)], dim=1)
            if taps is not None:
                taps["xvector.block%d" % bi] = x
            p = "xvector.transit%d" % bi
            x = F.conv1d(F.relu(_bn(sd, p + ".nonlinear.batchnorm", x)), _t(sd, p + ".linear.weight"))
            if taps is not None:
                taps[p] = x
        x = F.relu(_bn(sd, "xvector.out_nonlinear.batchnorm", x))
        stats = torch.cat([x.mean(-1), x.std(-1, unbiased=True)], dim=-1)
        if taps is not None:
            taps["xvector.stats"] = stats
        e = F.conv1d(stats.unsqueeze(-1), _t(sd, "xvector.dense.linear.weight")).squeeze(-1)
        e = _bn(sd, "xvector.dense.nonlinear.batchnorm", e, affine=False)
        return e
