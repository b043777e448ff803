"""numpy restatement of the trial metrics of the reference, speakerlab/utils/score_metrics.py:57-105
(compute_pmiss_pfa_rbst, compute_eer, compute_c_norm), and of the turn merging of
speakerlab/bin/infer_diarization.py:780-797 (compressed_seg).  TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import numpy as np


def pmiss_pfa(scores, labels):
    """score_metrics.py:57-77 with unit weights."""
    idx = np.argsort(scores, kind="stable")
    lab = labels[idx]
    tgt = (lab == 1).astype("f8")
    imp = (lab == 0).astype("f8")
    return np.cumsum(tgt) / tgt.sum(), 1 - np.cumsum(imp) / imp.sum()


def eer(fnr, fpr, scores):
    """score_metrics.py:80-94."""
    d = fnr - fpr
    x1 = np.flatnonzero(d >= 0)[0]
    x2 = np.flatnonzero(d < 0)[-1]
    a = (fnr[x1] - fpr[x1]) / (fpr[x2] - fpr[x1] - (fnr[x2] - fnr[x1]))
    return fnr[x1] + a * (fnr[x2] - fnr[x1]), np.sort(scores, kind="stable")[x1]


def c_norm(fnr, fpr, p_target, c_miss=1, c_fa=1):
    """score_metrics.py:97-105."""
    return min(c_miss * fnr * p_target + c_fa * fpr * (1 - p_target)) / min(c_miss * p_target, c_fa * (1 - p_target))


def compressed_seg(seg_list):
    """infer_diarization.py:780-797, sequential form: seg_list = [[st, ed, label], ...]."""
    out = []
    for i, (st, ed, c) in enumerate(seg_list):
        if i == 0:
            out.append([st, ed, c])
        elif c == out[-1][2]:
            if st > out[-1][1]:
                out.append([st, ed, c])
            else:
                out[-1][1] = ed
        else:
            if st < out[-1][1]:
                p = (out[-1][1] + st) / 2
                out[-1][1] = p
                st = p
            out.append([st, ed, c])
    return out
