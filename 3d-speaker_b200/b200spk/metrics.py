"""EER / minDCF of a trial list on the GPU: ``speakerlab/utils/score_metrics.py:57-105`` as called by
``speakerlab/bin/compute_score_metrics.py:128-135`` (compute_pmiss_pfa_rbst -> compute_eer -> compute_c_norm).

The scores of a million trials are already on the device (``cosine_pairs``); the DET curve is one sort plus two
prefix sums there, and only the three scalars travel back.  Sort and scan are ``torch.sort`` / ``torch.cumsum``
(device plumbing); the arithmetic follows the reference in float64, including its interpolation of the EER between
the two operating points around the crossing.
"""
import torch


def det_curve(scores, labels):
    """scores [N] float, labels [N] {0,1} (device tensors) -> (fnr, fpr, sorted scores), float64, ascending score."""
    order = torch.argsort(scores, stable=True)
    lab = labels[order].to(torch.float64)
    tgt = torch.cumsum(lab, 0)
    imp = torch.cumsum(1.0 - lab, 0)
    fnr = tgt / tgt[-1]
    fpr = 1.0 - imp / imp[-1]
    return fnr, fpr, scores[order]


def eer_min_dcf(scores, labels, p_target=0.01, c_miss=1.0, c_fa=1.0):
    """-> (eer, eer_threshold, min_dcf) as Python floats."""
    fnr, fpr, s = det_curve(scores, labels)
    diff = fnr - fpr
    x1 = int(torch.nonzero(diff >= 0)[0])              # first operating point with miss >= false alarm
    x2 = int(torch.nonzero(diff < 0)[-1])              # last one before the crossing
    a = (fnr[x1] - fpr[x1]) / (fpr[x2] - fpr[x1] - (fnr[x2] - fnr[x1]))
    eer = fnr[x1] + a * (fnr[x2] - fnr[x1])
    c_det = torch.min(c_miss * fnr * p_target + c_fa * fpr * (1.0 - p_target))
    c_def = min(c_miss * p_target, c_fa * (1.0 - p_target))
    return float(eer), float(s[x1]), float(c_det / c_def)
