"""numpy restatement of the reference spectral-clustering back end.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
speakerlab/process/cluster.py:23-112 (SpectralCluster) stage by stage; the eigensolve
and k-means stay calls into the same third-party code the reference calls
(scipy.sparse.linalg.eigsh, cluster.py:90; sklearn k_means, cluster.py:104 - container
scipy 1.18.1 / sklearn 1.9.0, reference pins scipy>=1.7.0 / scikit-learn==1.0.2).
Pinned by tests/golden/cluster_*.npz minted from the imported reference with
``np.random.seed`` set before each call.
"""
import numpy as np
import scipy.sparse.linalg
from sklearn.cluster._kmeans import k_means


def sim_mat(X):
    """cluster.py:59-62: sklearn cosine_similarity = normalize(X) @ normalize(X).T."""
    X = np.asarray(X)
    nrm = np.sqrt((X * X).sum(axis=1, keepdims=True))
    nrm[nrm == 0.0] = 1.0
    Xn = X / nrm
    return Xn @ Xn.T


def prune_count(n, pval, min_pnum=6):
    """cluster.py:67-68: number of smallest entries zeroed per row."""
    return min(int((1 - pval) * n), n - min_pnum)


def p_pruning(A, pval, min_pnum=6):
    """cluster.py:64-77 (in place): zero the n_elems smallest entries of every row."""
    n_elems = prune_count(A.shape[0], pval, min_pnum)
    for i in range(A.shape[0]):
        low = np.argsort(A[i, :])[:n_elems]
        A[i, low] = 0
    return A


def laplacian(M):
    """cluster.py:79-84: zero diag, D = sum |M|, L = D - M (unnormalised)."""
    M = M.copy()
    M[np.diag_indices(M.shape[0])] = 0
    D = np.diag(np.sum(np.abs(M), axis=1))
    return D - M


def eigen_gaps(vals):
    """cluster.py:107-112."""
    return [float(vals[i + 1]) - float(vals[i]) for i in range(len(vals) - 1)]


def spec_embs(L, min_num_spks, max_num_spks, k_oracle=None):
    """cluster.py:86-100."""
    lambdas, vecs = scipy.sparse.linalg.eigsh(L, k=min(max_num_spks + 1, L.shape[0]), which="SM")
    if k_oracle is not None:
        k = k_oracle
    else:
        gaps = eigen_gaps(lambdas[min_num_spks - 1:max_num_spks + 1])
        k = int(np.argmax(gaps)) + min_num_spks
    return vecs[:, :k], k, lambdas


def spectral_cluster(X, min_num_spks=1, max_num_spks=10, pval=0.02, min_pnum=6, oracle_num=None,
                     return_stages=False):
    """SpectralCluster.__call__ (cluster.py:35-57).  X [N,D] -> int labels [N]."""
    A = sim_mat(X)
    A = p_pruning(A, pval, min_pnum)
    sym = 0.5 * (A + A.T)
    L = laplacian(sym)
    emb, k, lambdas = spec_embs(L, min_num_spks, max_num_spks, oracle_num)
    _, labels, _ = k_means(emb, k)
    if return_stages:
        return labels, dict(laplacian=L, lambdas=lambdas, k=k, emb=emb)
    return labels


def match_labels(ref, got):
    """Hungarian matching of cluster ids; returns got relabelled into ref's ids (unmatched
    ids keep fresh numbers) so 'identical up to permutation' becomes array equality."""
    from scipy.optimize import linear_sum_assignment
    ref = np.asarray(ref)
    got = np.asarray(got)
    ru, gu = np.unique(ref), np.unique(got)
    cost = np.zeros((len(gu), len(ru)), dtype=np.int64)
    for i, g in enumerate(gu):
        for j, r in enumerate(ru):
            cost[i, j] = -np.sum((got == g) & (ref == r))
    gi, rj = linear_sum_assignment(cost)
    mapping = {gu[i]: ru[j] for i, j in zip(gi, rj)}
    nxt = int(max(ru.max(), gu.max())) + 1
    out = np.empty_like(got)
    for g in gu:
        if g not in mapping:
            mapping[g] = nxt
            nxt += 1
        out[got == g] = mapping[g]
    return out


# ---------------------------------------------------------------------------- AHC / CommonClustering
def ahc(X, fix_cos_thr=0.4):
    """AHCluster.__call__ (speakerlab/process/cluster.py:139-156) with scipy's average linkage standing in for
    fastcluster.linkage (same algorithm and output format; fastcluster is not installed here)."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform
    scr = sim_mat(np.asarray(X))
    scr = squareform(-scr, checks=False)
    lin = linkage(scr, method="average")
    adjust = abs(lin[:, 2].min())
    lin[:, 2] += adjust
    return fcluster(lin, -fix_cos_thr + adjust, criterion="distance") - 1


def filter_minor_cluster(labels, x, min_cluster_size):
    """CommonClustering.filter_minor_cluster (cluster.py:200-219)."""
    from sklearn.metrics.pairwise import cosine_similarity
    cset = np.unique(labels)
    csize = np.array([(labels == i).sum() for i in cset])
    minor_idx = np.where(csize <= min_cluster_size)[0]
    if len(minor_idx) == 0:
        return labels
    minor_cset = cset[minor_idx]
    major_idx = np.where(csize > min_cluster_size)[0]
    if len(major_idx) == 0:
        return np.zeros_like(labels)
    major_cset = cset[major_idx]
    major_center = np.stack([x[labels == i].mean(0) for i in major_cset])
    for i in range(len(labels)):
        if labels[i] in minor_cset:
            labels[i] = major_cset[cosine_similarity(x[i][np.newaxis], major_center).argmax()]
    return labels


def merge_by_cos(labels, x, cos_thr):
    """CommonClustering.merge_by_cos (cluster.py:221-238)."""
    from sklearn.metrics.pairwise import cosine_similarity
    while True:
        cset = np.unique(labels)
        if len(cset) == 1:
            break
        centers = np.stack([x[labels == i].mean(0) for i in cset])
        affinity = np.triu(cosine_similarity(centers, centers), 1)
        idx = np.unravel_index(np.argmax(affinity), affinity.shape)
        if affinity[idx] < cos_thr:
            break
        c1, c2 = cset[np.array(idx)]
        labels[labels == c2] = c1
    return labels


def common_clustering(X, cluster_type="spectral", cluster_line=40, mer_cos=None, min_cluster_size=4, **kw):
    """CommonClustering.__call__ (cluster.py:184-198)."""
    X = np.asarray(X)
    if X.shape[0] <= 1:
        return np.zeros(X.shape[0], dtype=int)
    if X.shape[0] < cluster_line or cluster_type == "AHC":
        labels = ahc(X, **(kw if cluster_type == "AHC" else {}))
    else:
        labels = spectral_cluster(X, **kw)
    labels = filter_minor_cluster(np.array(labels), X, min_cluster_size)
    if mer_cos is not None:
        labels = merge_by_cos(labels, X, mer_cos)
    return labels
