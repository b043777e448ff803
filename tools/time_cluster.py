"""Timing + convergence of the spectral back end on a synthetic [N, D] embedding set."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np, torch
import b200spk
from oracle import cluster_oracle, gen_golden
for n, d, k in ((1500, 192, 6), (4799, 192, 8)):
    X, truth = gen_golden.cluster_input(n, d, k, 23)
    sc = b200spk.SpectralCluster(min_num_spks=1, max_num_spks=15, pval=0.012)
    np.random.seed(0); sc(X)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    np.random.seed(0); lab = sc(X)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    np.random.seed(0); t2 = time.perf_counter(); ref, st = cluster_oracle.spectral_cluster(X, 1, 15, 0.012, return_stages=True); t3 = time.perf_counter()
    print("N=%d gpu %.1f ms (krylov %d, kmeans iters %d, k=%d)  cpu oracle %.1f ms  labels equal: %s  max|dlambda| %.2e" % (
        n, 1e3 * (t1 - t0), sc.last["krylov"], sc.last["kmeans_iters"], sc.last["k"], 1e3 * (t3 - t2),
        np.array_equal(cluster_oracle.match_labels(ref, lab), ref), np.abs(sc.last["lambdas"] - st["lambdas"]).max()))
