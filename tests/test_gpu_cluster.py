"""GPU: spectral-clustering back end through the C ABI vs the oracle and the golden vectors."""
import os

import numpy as np
import pytest
import torch

import b200spk
from oracle import cluster_oracle, gen_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "cluster.npz"))


@pytest.mark.parametrize("case", gen_golden.cluster_cases(), ids=lambda c: c[0])
def test_labels_and_spectrum_vs_golden(gold, case):
    name, n, d, k, seed, kw = case
    X, truth = gen_golden.cluster_input(n, d, k, seed)
    X0 = X.copy()
    sc = b200spk.SpectralCluster(**kw)
    np.random.seed(0)
    labels = sc(X)
    assert np.array_equal(X, X0)                                   # input untouched (cluster.py call sites reuse it)
    assert labels.shape == (n,) and labels.flags.writeable and np.issubdtype(labels.dtype, np.integer)
    assert sc.last["k"] == int(gold[name + ".k"])
    # eigenvalues: a near-tie at the pruning cut can keep a different (equally weighted) edge than
    # the reference's BLAS does, which moves bulk eigenvalues by O(1e-3); the count-deciding
    # eigengap is O(1).  Tight solver accuracy is asserted against a dense eigh of OUR Laplacian below.
    np.testing.assert_allclose(sc.last["lambdas"], gold[name + ".lambdas"], atol=5e-3)
    ref = gold[name + ".labels"]
    assert np.array_equal(cluster_oracle.match_labels(ref, labels), ref)       # identical up to permutation


@pytest.mark.parametrize("case", gen_golden.cluster_cases(), ids=lambda c: c[0])
def test_laplacian_stage_vs_oracle(case):
    name, n, d, k, seed, kw = case
    X, _ = gen_golden.cluster_input(n, d, k, seed)
    sc = b200spk.SpectralCluster(**kw)
    lap = sc.laplacian(torch.from_numpy(X).cuda())[:n, :n].cpu().numpy()
    A = cluster_oracle.sim_mat(X)
    P = cluster_oracle.p_pruning(A.copy(), sc.pval, sc.min_pnum)
    Lref = cluster_oracle.laplacian(0.5 * (P + P.T))
    keep = n - cluster_oracle.prune_count(n, sc.pval, sc.min_pnum)
    assert np.all((P != 0).sum(1) == keep)
    # the same edges survive the pruning (fp32-accurate affinity) except for near-ties at the cut:
    # the two affinities differ by ~1e-7, so a row whose keep-th and (keep+1)-th values are that
    # close may keep the other one.  Allow at most 1 such row per 200.
    diff_rows = np.unique(np.nonzero((lap != 0) != (Lref != 0))[0])
    assert len(diff_rows) <= max(1, n // 200), len(diff_rows)
    same = (lap != 0) == (Lref != 0)
    np.testing.assert_allclose(lap[same & ~np.eye(n, dtype=bool)], Lref[same & ~np.eye(n, dtype=bool)], rtol=5e-6, atol=2e-6)
    ok_rows = np.setdiff1d(np.arange(n), np.unique(np.concatenate([diff_rows, np.nonzero((lap != 0) != (Lref != 0))[1]])))
    np.testing.assert_allclose(np.diag(lap)[ok_rows], np.diag(Lref)[ok_rows], rtol=5e-6)


@pytest.mark.parametrize("case", gen_golden.cluster_cases(), ids=lambda c: c[0])
def test_eigensolver_accuracy_on_own_laplacian(case):
    name, n, d, k, seed, kw = case
    X, _ = gen_golden.cluster_input(n, d, k, seed)
    sc = b200spk.SpectralCluster(**kw)
    lap_d = sc.laplacian(torch.from_numpy(X).cuda())
    kk = min(sc.max_num_spks + 1, n)
    lam, vec = sc.eig_smallest(lap_d, n, kk)
    w = np.linalg.eigvalsh(lap_d[:n, :n].cpu().numpy().astype(np.float64))[:kk]
    np.testing.assert_allclose(lam, w, atol=5e-4)


def test_eigensolver_vs_dense_eigh():
    X, _ = gen_golden.cluster_input(900, 64, 5, 31)
    sc = b200spk.SpectralCluster(max_num_spks=15, pval=0.012)
    lap_d = sc.laplacian(torch.from_numpy(X).cuda())
    lam, vec = sc.eig_smallest(lap_d, 900, 16)
    lap = lap_d[:900, :900].cpu().numpy().astype(np.float64)
    w, v = np.linalg.eigh(lap)
    np.testing.assert_allclose(lam, w[:16], atol=5e-4)
    # leading invariant subspace (below the eigengap): principal angles ~ 0
    q = int(np.argmax(np.diff(w[:16]))) + 1
    Q, _ = np.linalg.qr(vec.cpu().numpy()[:, :q].astype(np.float64))
    s = np.linalg.svd(Q.T @ v[:, :q], compute_uv=False)
    assert s.min() > 1 - 1e-4


def test_pruning_ties_and_duplicates():
    # exact duplicate rows produce exact ties in the affinity: still exactly `keep` survivors per row
    rng = np.random.default_rng(5)
    base = rng.standard_normal((40, 32)).astype(np.float32)
    X = np.concatenate([base, base, base, rng.standard_normal((80, 32)).astype(np.float32)])
    n = X.shape[0]
    sc = b200spk.SpectralCluster(pval=0.1)
    n_elems = min(int((1 - 0.1) * n), n - 6)
    lap = sc.laplacian(torch.from_numpy(X).cuda())[:n, :n].cpu().numpy()
    assert np.allclose(lap, lap.T, atol=1e-6)
    assert np.all(np.diag(lap) >= 0)
    assert np.allclose(lap.sum(1), 0, atol=1e-4)            # unnormalised Laplacian: rows sum to zero (entries >= 0 here)
    assert n - n_elems >= 6


def test_oracle_num_and_pval_kwargs():
    X, truth = gen_golden.cluster_input(400, 64, 3, 41)
    sc = b200spk.SpectralCluster(min_num_spks=1, max_num_spks=15, pval=0.012)
    np.random.seed(0)
    a = sc(X, speaker_num=3)
    assert len(np.unique(a)) == 3
    assert np.array_equal(cluster_oracle.match_labels(truth, a), truth)
    np.random.seed(0)
    b = sc(X, pval=0.05)
    assert np.array_equal(cluster_oracle.match_labels(truth, b), truth)


def test_small_n_rejected_like_eigsh():
    sc = b200spk.SpectralCluster(max_num_spks=15)
    with pytest.raises(ValueError):
        sc(np.random.default_rng(0).standard_normal((10, 16)).astype(np.float32))


def test_cosine_pairs():
    rng = np.random.default_rng(9)
    E = rng.standard_normal((500, 192)).astype(np.float32)
    a = rng.integers(0, 500, 4000).astype(np.int32)
    b = rng.integers(0, 500, 4000).astype(np.int32)
    got = b200spk.cosine_pairs(torch.from_numpy(E).cuda(), torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    ref = (E[a] * E[b]).sum(1) / (np.linalg.norm(E[a], axis=1) * np.linalg.norm(E[b], axis=1))
    np.testing.assert_allclose(got, ref, atol=2e-6)


def test_self_contained_c_eigensolver_matches_split_path():
    # spk_eig_smallest (built-in QL) and the spk_lanczos_extend/ritz + LAPACK path share the Krylov run
    import ctypes as C
    from b200spk import _lib
    X, _ = gen_golden.cluster_input(500, 64, 4, 51)
    sc = b200spk.SpectralCluster(max_num_spks=15, pval=0.012)
    lap = sc.laplacian(torch.from_numpy(X).cuda())
    lam_a, vec_a = sc.eig_smallest(lap, 500, 16)
    L = _lib.lib()
    ws = torch.empty(int(L.spk_eig_workspace_bytes(500, 16)), dtype=torch.uint8, device="cuda")
    lam_b = np.empty(16, dtype=np.float32)
    vec_b = torch.empty((500, 16), dtype=torch.float32, device="cuda")
    m = _lib.check(L.spk_eig_smallest(C.c_void_p(lap.data_ptr()), 500, 16, lam_b.ctypes.data_as(C.c_void_p),
                                      C.c_void_p(vec_b.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), None))
    assert m > 0
    np.testing.assert_allclose(lam_a, lam_b, atol=5e-4)
