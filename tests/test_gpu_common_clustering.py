"""GPU: AHCluster / CommonClustering mirrors (speakerlab/process/cluster.py:139-239) vs golden labels and the oracle."""
import os

import numpy as np
import pytest

import b200spk
from oracle import cluster_oracle, gen_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "common_clustering.npz"))


@pytest.mark.parametrize("case", gen_golden.common_cases(), ids=lambda c: c[0])
def test_common_clustering_vs_golden(gold, case):
    name, n, d, k, seed, noise, outliers, kw = case
    X, _ = gen_golden.common_input(n, d, k, seed, noise, outliers)
    keep = X.copy()
    np.random.seed(0)
    got = b200spk.CommonClustering(**kw)(X)
    assert np.array_equal(X, keep)                       # the caller's embeddings are left alone
    ref = gold[name + ".labels"]
    assert got.shape == ref.shape and got.flags.writeable
    assert np.array_equal(cluster_oracle.match_labels(ref, got), ref)


@pytest.mark.parametrize("case", [c for c in gen_golden.common_cases() if c[7]["cluster_type"] == "AHC"], ids=lambda c: c[0])
def test_ahc_raw_labels_vs_golden(gold, case):
    name, n, d, k, seed, noise, outliers, kw = case
    X, _ = gen_golden.common_input(n, d, k, seed, noise, outliers)
    got = b200spk.AHCluster(kw.get("fix_cos_thr", 0.4))(X)
    ref = gold[name + ".raw_ahc"]
    assert np.array_equal(cluster_oracle.match_labels(ref, got), ref)
    # clusters are numbered by their smallest member: the first point is always in cluster 0
    assert got[0] == 0 and set(np.unique(got)) == set(range(int(got.max()) + 1))


@pytest.mark.parametrize("n,thr", [(1, 0.4), (2, 0.4), (2, -1.0), (3000, 0.3)])
def test_ahc_edge_sizes_vs_oracle(n, thr):
    """Single point, a pair below and above the cut, and a size where the per-row neighbour refresh lists overflow
    into the full rescan path; thresholds where the cut is unambiguous."""
    rng = np.random.default_rng(n)
    centers = rng.standard_normal((4, 48))
    X = (centers[rng.integers(0, 4, n)] + 0.3 * rng.standard_normal((n, 48))).astype(np.float32)
    got = b200spk.AHCluster(thr)(X)
    if n == 1:
        assert got.tolist() == [0]
        return
    ref = cluster_oracle.ahc(X, thr)
    assert np.array_equal(cluster_oracle.match_labels(ref, got), ref)


def test_short_recordings_take_ahc():
    """Fewer than cluster_line segments -> AHC with the default threshold even when spectral is configured
    (cluster.py:178-181, 189-190)."""
    X, _ = gen_golden.common_input(25, 64, 2, 33, 0.35, 0)
    cc = b200spk.CommonClustering("spectral", min_num_spks=1, max_num_spks=15, pval=0.012)
    got = cc(X)
    ref = cluster_oracle.ahc(X, 0.4)
    assert np.array_equal(cluster_oracle.match_labels(ref, got), ref)
    assert b200spk.CommonClustering("AHC")(X[:1]).tolist() == [0]
    with pytest.raises(ValueError):
        b200spk.CommonClustering("umap_hdbscan")
