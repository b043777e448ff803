"""Label post-processing after clustering: sub-segment labels -> speaker turns -> RTTM.

Mirrors ``compressed_seg`` (speakerlab/bin/infer_diarization.py:780-797) and ``make_rttms``
(egs/3dspeaker/speaker-diarization/local/cluster_and_postprocess.py:25-50): consecutive sub-segments with the same
label merge into one turn (unless separated by a gap), and where two different speakers' segments overlap the
boundary moves to the middle of the overlap.  Host side: it runs once per recording on a few thousand (start, end,
label) triples, after the labels have come back from the GPU.

Implemented on run boundaries instead of a per-segment scan: for sub-segments with non-decreasing start and end
times (what ``chunk()`` produces for every VAD region) a turn is a maximal run of equal labels without a gap, its end
is the last member's end, and the overlap rule only ever applies between neighbouring runs.
"""
import numpy as np


def compress_segments(segments, labels):
    """segments: [[start, end], ...] seconds (sorted); labels: int [N].  Returns a list of [start, end, label]."""
    seg = np.asarray(segments, dtype=np.float64).reshape(-1, 2)
    lab = np.asarray(labels).reshape(-1)
    assert seg.shape[0] == lab.shape[0]
    if seg.shape[0] == 0:
        return []
    st, ed = seg[:, 0], seg[:, 1]
    assert np.all(np.diff(st) >= 0) and np.all(np.diff(ed) >= 0), "sub-segments must be sorted by time"
    # a new turn starts where the label changes, or where a same-label segment starts after the previous one ended
    new_turn = np.ones(len(lab), dtype=bool)
    new_turn[1:] = (lab[1:] != lab[:-1]) | (st[1:] > ed[:-1])
    first = np.flatnonzero(new_turn)
    last = np.append(first[1:], len(lab)) - 1
    t_st, t_ed, t_lab = st[first].copy(), ed[last].copy(), lab[first]
    # neighbouring turns of different speakers that overlap meet in the middle of the overlap
    if len(first) > 1:
        overlap = (t_lab[1:] != t_lab[:-1]) & (t_st[1:] < t_ed[:-1])
        mid = 0.5 * (t_ed[:-1] + t_st[1:])
        t_ed[:-1] = np.where(overlap, mid, t_ed[:-1])
        t_st[1:] = np.where(overlap, mid, t_st[1:])
    return [[float(a), float(b), int(c)] for a, b, c in zip(t_st, t_ed, t_lab)]


def rttm_lines(turns, rec_id):
    """RTTM text lines with 1-based speaker ids, the format of make_rttms (cluster_and_postprocess.py:45-50)."""
    return ["SPEAKER {} 0 {:.3f} {:.3f} <NA> <NA> {:d} <NA> <NA>\n".format(rec_id, s, e - s, int(c) + 1) for s, e, c in turns]


def write_rttm(path, segments, labels, rec_id):
    turns = compress_segments(segments, labels)
    with open(path, "w") as f:
        f.writelines(rttm_lines(turns, rec_id))
    return turns
