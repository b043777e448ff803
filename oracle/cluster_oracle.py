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
