"""Plain-PyTorch fp32 CPU restatement of the ECAPA-TDNN forward, driven by a state_dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional, eval-mode BatchNorm, written against
the reference's state_dict key names.  Follows:

  ECAPA_TDNN.forward                  speakerlab/models/ecapa_tdnn/ECAPA_TDNN.py:430-463
  Conv1d ('same' reflect padding)     :42-106, get_padding_elem :29-39
  TDNNBlock.forward (conv-ReLU-BN)    :150-151
  Res2NetBlock.forward                :180-191
  SEBlock.forward                     :209-222
  AttentiveStatisticsPooling.forward  :243-287
  SERes2NetBlock.forward              :336-347

Pinned by tests/golden/ecapa.npz (minted from the imported reference).
"""
import torch
import torch.nn.functional as F


def _t(sd, key):
    v = sd[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(v)


def conv1d_same(sd, p, x, dilation=1):
    w, b = _t(sd, p + ".conv.weight"), _t(sd, p + ".conv.bias")
    k = w.shape[2]
    pad = dilation * (k - 1) // 2
    if pad:
        x = F.pad(x, (pad, pad), mode="reflect")
    return F.conv1d(x, w, b, dilation=dilation)


def tdnn(sd, p, x, dilation=1):
    y = torch.relu(conv1d_same(sd, p + ".conv", x, dilation))
    q = p + ".norm.norm"
    return F.batch_norm(y, _t(sd, q + ".running_mean"), _t(sd, q + ".running_var"), _t(sd, q + ".weight"), _t(sd, q + ".bias"),
                        training=False, eps=1e-5)


def res2net(sd, p, x, scale, dilation):
    ys = []
    y = None
    for i, xi in enumerate(torch.chunk(x, scale, dim=1)):
        if i == 0:
            y = xi
        elif i == 1:
            y = tdnn(sd, "%s.blocks.%d" % (p, i - 1), xi, dilation)
        else:
            y = tdnn(sd, "%s.blocks.%d" % (p, i - 1), xi + y, dilation)
        ys.append(y)
    return torch.cat(ys, dim=1)


def se_block(sd, p, x):
    s = x.mean(dim=2, keepdim=True)
    s = torch.relu(conv1d_same(sd, p + ".conv1", s))
    s = torch.sigmoid(conv1d_same(sd, p + ".conv2", s))
    return s * x


def se_res2net(sd, p, x, scale, dilation):
    residual = x
    if (p + ".shortcut.conv.weight") in sd:
        residual = conv1d_same(sd, p + ".shortcut", x)
    x = tdnn(sd, p + ".tdnn1", x)
    x = res2net(sd, p + ".res2net_block", x, scale, dilation)
    x = tdnn(sd, p + ".tdnn2", x)
    x = se_block(sd, p + ".se_block", x)
    return x + residual


def asp(sd, p, x, eps=1e-12):
    L = x.shape[-1]

    def stats(x, m):
        mean = (m * x).sum(2)
        std = torch.sqrt((m * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(eps))
        return mean, std

    m = torch.full((x.shape[0], 1, L), 1.0 / L)
    mean, std = stats(x, m)
    attn = torch.cat([x, mean.unsqueeze(2).repeat(1, 1, L), std.unsqueeze(2).repeat(1, 1, L)], dim=1)
    attn = conv1d_same(sd, p + ".conv", torch.tanh(tdnn(sd, p + ".tdnn", attn)))
    attn = F.softmax(attn, dim=2)
    mean, std = stats(x, attn)
    return torch.cat((mean, std), dim=1).unsqueeze(2)


def forward(sd, feats, dilations=(1, 2, 3, 4, 1), scale=8, taps=None):
    """feats [B, T, F] float32 -> embeddings [B, lin_neurons]."""
    x = torch.as_tensor(feats, dtype=torch.float32).transpose(1, 2)
    n_blocks = 1 + sum(1 for k in sd if k.endswith(".tdnn1.conv.conv.weight"))
    xl = []
    with torch.no_grad():
        x = tdnn(sd, "blocks.0", x, dilations[0])
        xl.append(x)
        for i in range(1, n_blocks):
            x = se_res2net(sd, "blocks.%d" % i, x, scale, dilations[i])
            xl.append(x)
            if taps is not None:
                taps["blocks.%d" % i] = x
        x = tdnn(sd, "mfa", torch.cat(xl[1:], dim=1), dilations[-1])
        if taps is not None:
            taps["mfa"] = x
        x = asp(sd, "asp", x)
        q = "asp_bn.norm"
        x = F.batch_norm(x, _t(sd, q + ".running_mean"), _t(sd, q + ".running_var"), _t(sd, q + ".weight"), _t(sd, q + ".bias"),
                         training=False, eps=1e-5)
        x = conv1d_same(sd, "fc", x)
    return x.transpose(1, 2).squeeze(1)
