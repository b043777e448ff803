"""Drop-in for ``speakerlab.process.cluster.SpectralCluster`` (speakerlab/process/cluster.py:23-112).

Same constructor and call convention (numpy ``X [N, D]`` in, writable numpy int labels out,
``pval=`` / ``speaker_num=`` keyword overrides, ``X`` is never mutated).  The arithmetic runs on
the GPU through the C ABI: cosine affinity (tensor-core GEMM at fp32 accuracy), p-pruning,
symmetrised unnormalised Laplacian, smallest eigenpairs (Lanczos) and Lloyd k-means.  Host side:
the eigengap rule (cluster.py:93-96, 107-112) and the k-means++ seeding, which draws from numpy's
GLOBAL RandomState in the same order sklearn's ``k_means(emb, k)`` does (sklearn
``_kmeans_plusplus``), so ``np.random.seed(s)`` before the call pins the result exactly as it does
for the reference.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib


def _kmeans_plusplus(X, n_clusters, random_state):
    """Greedy k-means++ seeding, sklearn/cluster/_kmeans.py:_kmeans_plusplus (uniform sample weights):
    same RandomState calls in the same order."""
    n_samples = X.shape[0]
    centers = np.empty((n_clusters, X.shape[1]), dtype=X.dtype)
    n_local_trials = 2 + int(np.log(n_clusters))
    sample_weight = np.ones(n_samples, dtype=X.dtype)
    center_id = random_state.choice(n_samples, p=sample_weight / sample_weight.sum())
    centers[0] = X[center_id]
    X64 = X.astype(np.float64)
    x_sq = (X64 * X64).sum(axis=1)

    def dist_sq(Y):
        # sklearn upcasts float32 inputs to float64 for this product (_euclidean_distances_upcast)
        Y64 = Y.astype(np.float64)
        d = x_sq[None, :] - 2.0 * (Y64 @ X64.T) + (Y64 * Y64).sum(axis=1)[:, None]
        np.maximum(d, 0, out=d)
        return d.astype(X.dtype)

    closest = dist_sq(centers[0:1])
    current_pot = closest @ sample_weight
    for c in range(1, n_clusters):
        rand_vals = random_state.uniform(size=n_local_trials) * current_pot
        candidate_ids = np.searchsorted(np.cumsum(sample_weight * closest), rand_vals)
        np.clip(candidate_ids, None, closest.size - 1, out=candidate_ids)
        d2c = dist_sq(X[candidate_ids])
        np.minimum(closest, d2c, out=d2c)
        pots = d2c @ sample_weight.reshape(-1, 1)
        best = int(np.argmin(pots))
        current_pot = pots[best]
        closest = d2c[best:best + 1]
        centers[c] = X[candidate_ids[best]]
    return centers


class SpectralCluster:
    """A spectral clustering method using the unnormalised Laplacian of the pruned cosine affinity."""

    def __init__(self, min_num_spks=1, max_num_spks=10, pval=0.02, min_pnum=6, oracle_num=None, device="cuda:0"):
        self.min_num_spks = min_num_spks
        self.max_num_spks = max_num_spks
        self.min_pnum = min_pnum
        self.pval = pval
        self.k = oracle_num
        self.device = torch.device(device)
        self.last = {}          # stage results of the last call (eigenvalues, k, Krylov size) for inspection

    # ------------------------------------------------------------------ stages (device)
    def laplacian(self, X, pval=None):
        """X: torch f32 [N, D] on the device -> (L [Np, Np] with row pitch Np, N)."""
        L = _lib.lib()
        N, D = X.shape
        pval = self.pval if pval is None else pval
        n_elems = min(int((1 - pval) * N), N - self.min_pnum)        # cluster.py:67-68
        keep = N - max(n_elems, 0)
        Np = (N + 15) // 16 * 16
        lap = torch.empty((Np, Np), dtype=torch.float32, device=X.device)
        ws = torch.empty(int(L.spk_affinity_workspace_bytes(N, D)), dtype=torch.uint8, device=X.device)
        with torch.cuda.device(X.device):
            _lib.check(L.spk_affinity_laplacian(C.c_void_p(X.data_ptr()), N, D, keep, C.c_void_p(lap.data_ptr()),
                                                C.c_void_p(ws.data_ptr()), ws.numel(), _lib.current_stream_ptr()))
        return lap

    def eig_smallest(self, lap, N, k, tol=1e-4):
        """k smallest eigenpairs of the Laplacian: Lanczos on the GPU (spk_lanczos_extend), the small
        tridiagonal problem by LAPACK on the host, Ritz vectors on the GPU (spk_lanczos_ritz).
        Stops when every wanted Ritz pair has residual estimate |beta_m s_m| <= tol."""
        from scipy.linalg import eigh_tridiagonal
        L = _lib.lib()
        ws = torch.empty(int(L.spk_eig_workspace_bytes(N, k)), dtype=torch.uint8, device=lap.device)
        m_max = int(L.spk_lanczos_max_dim(N, k))
        alpha = np.zeros(m_max + 1, dtype=np.float32)
        beta = np.zeros(m_max + 1, dtype=np.float32)
        sigma = C.c_float(0.0)
        m_done, m_to = 0, min(m_max, max(8 * k, 96))
        with torch.cuda.device(lap.device):
            while True:
                _lib.check(L.spk_lanczos_extend(C.c_void_p(lap.data_ptr()), N, k, m_done, m_to,
                                                alpha.ctypes.data_as(C.c_void_p), beta.ctypes.data_as(C.c_void_p),
                                                C.byref(sigma), C.c_void_p(ws.data_ptr()), ws.numel(),
                                                _lib.current_stream_ptr()))
                m_done = m_to
                d = alpha[:m_done].astype(np.float64)
                e = beta[:m_done - 1].astype(np.float64)
                theta, S = eigh_tridiagonal(d, e, select="i", select_range=(m_done - k, m_done - 1))
                resid = np.abs(float(beta[m_done - 1]) * S[-1, :])
                if resid.max() <= tol or m_done >= m_max or m_done >= N:
                    break
                m_to = min(m_max, m_done + max(6 * k, 96))
            # largest Ritz values of sigma*I - L first == smallest eigenvalues of L first
            theta, S = theta[::-1], np.ascontiguousarray(S[:, ::-1], dtype=np.float32)
            evals = (float(sigma.value) - theta).astype(np.float32)
            evecs = torch.empty((N, k), dtype=torch.float32, device=lap.device)
            _lib.check(L.spk_lanczos_ritz(N, k, m_done, S.ctypes.data_as(C.c_void_p), C.c_void_p(evecs.data_ptr()),
                                          C.c_void_p(ws.data_ptr()), ws.numel(), _lib.current_stream_ptr()))
        self.last["krylov"] = int(m_done)
        self.last["ritz_residual"] = float(resid.max())
        return evals, evecs

    def kmeans(self, emb, k):
        """emb: torch f32 [N, d] on the device -> numpy int32 labels.  Mirrors sklearn k_means(emb, k):
        mean-centring, tol = mean(var)*1e-4, k-means++ from the global numpy RNG, one init, Lloyd."""
        L = _lib.lib()
        N, d = emb.shape
        host = emb.cpu().numpy().astype(np.float32)
        host = host - host.mean(axis=0)
        tol = float(np.mean(np.var(host, axis=0)) * 1e-4)
        centers = _kmeans_plusplus(host, k, np.random.mtrand._rand).astype(np.float32)
        pts = torch.from_numpy(host).to(emb.device)
        labels = torch.empty(N, dtype=torch.int32, device=emb.device)
        ws = torch.empty(int(L.spk_kmeans_workspace_bytes(N, d, k)), dtype=torch.uint8, device=emb.device)
        inertia = C.c_float(0.0)
        with torch.cuda.device(emb.device):
            it = _lib.check(L.spk_kmeans(C.c_void_p(pts.data_ptr()), N, d, k, centers.ctypes.data_as(C.c_void_p), 300,
                                         C.c_float(tol), C.c_void_p(labels.data_ptr()), C.byref(inertia),
                                         C.c_void_p(ws.data_ptr()), ws.numel(), _lib.current_stream_ptr()))
        self.last["kmeans_iters"] = int(it)
        return labels.cpu().numpy()

    # ------------------------------------------------------------------ reference call convention
    def __call__(self, X, **kwargs):
        pval = kwargs.get('pval', None)
        oracle_num = kwargs.get('speaker_num', None)
        if isinstance(X, torch.Tensor):
            Xd = X.detach().to(self.device, torch.float32).contiguous()
        else:
            Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(self.device)
        N = Xd.shape[0]
        k_eig = min(self.max_num_spks + 1, N)
        if k_eig >= N:
            raise ValueError("k=%d must be less than N=%d (scipy eigsh would raise too)" % (k_eig, N))
        if k_eig > 32:
            raise ValueError("max_num_spks > 31 is not supported by the GPU eigensolver")
        t0 = time.perf_counter()
        lap = self.laplacian(Xd, pval)
        lambdas, vecs = self.eig_smallest(lap, N, k_eig)
        t1 = time.perf_counter()
        k_oracle = self.k if oracle_num is None else oracle_num
        if k_oracle is not None:
            num_of_spk = int(k_oracle)
        else:
            lam = lambdas[self.min_num_spks - 1:self.max_num_spks + 1]
            gaps = [float(lam[i + 1]) - float(lam[i]) for i in range(len(lam) - 1)]
            num_of_spk = int(np.argmax(gaps)) + self.min_num_spks
        self.last.update(lambdas=lambdas, k=num_of_spk)
        emb = vecs[:, :num_of_spk].contiguous()
        labels = self.kmeans(emb, num_of_spk)
        self.last["stages_s"] = {"affinity_laplacian_eig": t1 - t0, "kmeans": time.perf_counter() - t1}
        return labels


class AHCluster:
    """Drop-in for ``speakerlab.process.cluster.AHCluster`` (cluster.py:139-156): average-linkage agglomerative
    clustering on -cosine, cut at ``fix_cos_thr``.  The affinity GEMM and the whole merge loop run on the GPU
    (``spk_ahc``); labels come back numbered by the smallest member index of each cluster (the reference's numbering
    follows scipy's dendrogram order - callers only use label identity)."""

    def __init__(self, fix_cos_thr=0.4, device="cuda:0"):
        self.fix_cos_thr = fix_cos_thr
        self.device = torch.device(device)
        self.last = {}

    def __call__(self, X, **kwargs):
        if isinstance(X, torch.Tensor):
            Xd = X.detach().to(self.device, torch.float32).contiguous()
        else:
            Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(self.device)
        N, D = Xd.shape
        L = _lib.lib()
        labels = torch.empty(N, dtype=torch.int32, device=self.device)
        ws = torch.empty(int(L.spk_ahc_workspace_bytes(N, D)), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            k = _lib.check(L.spk_ahc(C.c_void_p(Xd.data_ptr()), N, D, C.c_float(self.fix_cos_thr), C.c_void_p(labels.data_ptr()),
                                     C.c_void_p(ws.data_ptr()), ws.numel(), _lib.current_stream_ptr()))
        self.last["k"] = int(k)
        return labels.cpu().numpy().astype(np.int64)


def _unit_rows(x):
    """Rows scaled to unit L2 norm in float64 (all-zero rows stay zero, as sklearn's normalize leaves them)."""
    x = np.asarray(x, dtype=np.float64)
    n = np.linalg.norm(x, axis=1, keepdims=True)
    return np.divide(x, n, out=np.zeros_like(x), where=n > 0)


class _ClusterStats:
    """Per-cluster running sums over the embeddings: centroids come from (sum, count), so reassigning or merging
    clusters never rescans the N x D matrix."""

    def __init__(self, labels, x):
        self.ids, self.member, self.count = np.unique(labels, return_inverse=True, return_counts=True)
        self.total = np.zeros((len(self.ids), x.shape[1]), dtype=np.float64)
        np.add.at(self.total, self.member, x.astype(np.float64, copy=False))

    def centroids(self, rows=None):
        rows = slice(None) if rows is None else rows
        return self.total[rows] / self.count[rows, None]


class CommonClustering:
    """Drop-in for ``speakerlab.process.cluster.CommonClustering`` (cluster.py:158-239): dispatch to the GPU back end
    (``spectral`` or ``AHC``; recordings shorter than ``cluster_line`` segments always take AHC), then the two label
    post-steps of the reference on the host (they touch K <= max_num_spks centroids, not the data path):

    * ``filter_minor_cluster`` (cluster.py:204-222): every member of a cluster with at most ``min_cluster_size``
      segments moves to the major cluster whose centroid (computed once, before any move) it is most cosine-similar
      to; if no cluster is major everything becomes label 0.  Done here as one [n_minor_rows, K_major] product.
    * ``merge_by_cos`` (cluster.py:224-239): repeatedly merge the most similar centroid pair while its cosine is
      >= ``mer_cos``; the pair keeps the smaller label.  Done here on per-cluster (sum, count) statistics, so a
      merge is two row additions instead of a rescan of the embeddings."""

    def __init__(self, cluster_type, cluster_line=40, mer_cos=None, min_cluster_size=4, device="cuda:0", **kwargs):
        self.cluster_type = cluster_type
        self.cluster_line = cluster_line
        self.min_cluster_size = min_cluster_size
        self.mer_cos = mer_cos
        if cluster_type == 'spectral':
            self.cluster = SpectralCluster(device=device, **kwargs)
        elif cluster_type == 'AHC':
            self.cluster = AHCluster(device=device, **kwargs)
        else:
            raise ValueError('%s is not currently supported.' % cluster_type)
        self.cluster_for_short = self.cluster if cluster_type == 'AHC' else AHCluster(device=device)

    def __call__(self, X, **kwargs):
        assert len(X.shape) == 2, 'Shape of input should be [N, C]'
        n = X.shape[0]
        if n <= 1:
            return np.zeros(n, dtype=int)
        emb = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
        backend = self.cluster_for_short if n < self.cluster_line else self.cluster
        labels = np.array(backend(emb, **(kwargs if backend is self.cluster else {})), dtype=np.int64)
        labels = self.filter_minor_cluster(labels, emb, self.min_cluster_size)
        if self.mer_cos is not None:
            labels = self.merge_by_cos(labels, emb, self.mer_cos)
        return labels

    def filter_minor_cluster(self, labels, x, min_cluster_size):
        # like the reference, the instance attribute decides the size limit (the argument is ignored there too)
        st = _ClusterStats(labels, x)
        is_major = st.count > self.min_cluster_size
        if is_major.all():
            return labels
        if not is_major.any():
            return np.zeros_like(labels)
        movers = np.flatnonzero(~is_major[st.member])
        sim = _unit_rows(x[movers]) @ _unit_rows(st.centroids(is_major)).T
        labels[movers] = st.ids[is_major][sim.argmax(axis=1)]
        return labels

    def merge_by_cos(self, labels, x, cos_thr):
        assert cos_thr > 0 and cos_thr <= 1
        st = _ClusterStats(labels, x)
        ids, total, count = list(st.ids), st.total, st.count.astype(np.float64)
        while len(ids) > 1:
            c = _unit_rows(total / count[:, None])
            sim = np.triu(c @ c.T, k=1)
            a, b = divmod(int(sim.argmax()), len(ids))       # first maximum in row-major order, a < b
            if sim[a, b] < cos_thr:
                break
            labels[labels == ids[b]] = ids[a]
            total[a] += total[b]
            count[a] += count[b]
            keep = np.arange(len(ids)) != b
            total, count = total[keep], count[keep]
            del ids[b]
        return labels


def cosine_pairs(E, a, b):
    """out[i] = cos(E[a[i]], E[b[i]]): trial scoring (speakerlab/bin/compute_score_metrics.py:113-114).
    E torch f32 [N, D] on the device, a/b int32 index tensors on the device."""
    out = torch.empty(a.shape[0], dtype=torch.float32, device=E.device)
    with torch.cuda.device(E.device):
        _lib.check(_lib.lib().spk_cosine_pairs(C.c_void_p(E.data_ptr()), E.shape[0], E.shape[1], C.c_void_p(a.data_ptr()),
                                               C.c_void_p(b.data_ptr()), a.shape[0], C.c_void_p(out.data_ptr()),
                                               _lib.current_stream_ptr()))
    return out
