"""Mint golden vectors from the IMPORTED reference (runs only where /root/reference exists).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage (from the repo root):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden

The reference is imported from /root/reference (never copied); fastcluster/umap/hdbscan
are stubbed in sys.modules because speakerlab/process/cluster.py:10-20 imports them at
module scope and SpectralCluster uses none of them.  Inputs and weights are NOT stored:
they are regenerated from seeds by oracle/synth.py on both sides.  Outputs go to
tests/golden/*.npz together with the library versions that produced them.
"""
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("SPK_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("fastcluster", "umap", "hdbscan"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from speakerlab.process.processor import FBank
    from speakerlab.models.campplus.DTDNN import CAMPPlus
    from speakerlab.process.cluster import SpectralCluster
    return FBank, CAMPPlus, SpectralCluster


def import_eres2netv2():
    from speakerlab.models.eres2net.ERes2NetV2 import ERes2NetV2
    return ERes2NetV2


def versions():
    import scipy, sklearn, torch, torchaudio
    return json.dumps(dict(numpy=np.__version__, torch=torch.__version__, torchaudio=torchaudio.__version__,
                           scipy=scipy.__version__, sklearn=sklearn.__version__))


# the input cases; tests regenerate them with the same calls
def fbank_cases():
    from oracle import synth
    wav, _ = synth.fm_meeting(12.0, 3, seed=11)
    cases = {
        "noise_1p5s": synth.white_noise(2, 24000, seed=1),
        "noise_3s": synth.white_noise(1, 48000, seed=2),
        "fm_1p5s": synth.cut_windows(wav, synth.chunk(0.0, 12.0)[:3]),
        "one_frame": synth.white_noise(1, 400, seed=3),
        "ragged_559": synth.white_noise(1, 559, seed=4),
        "ragged_560": synth.white_noise(1, 560, seed=5),
        "loud": synth.white_noise(1, 8000, seed=6, scale=0.9),
        "silence_tail": np.concatenate([synth.white_noise(1, 4000, seed=7), np.zeros((1, 4000), np.float32)], axis=1),
    }
    return cases


def campplus_cases():
    """(name, embedding_size, batch, n_samples, weight seed, randomize_bn)"""
    return [
        ("e192_t148_bnrand", 192, 3, 24000, 101, True),
        ("e512_t148_bnrand", 512, 2, 24000, 102, True),
        ("e192_t298_bnrand", 192, 2, 48000, 103, True),   # T'=149 -> 2 seg-pooling windows
        ("e192_t148_freshbn", 192, 2, 24000, 104, False),
    ]


def campplus_input(batch, n_samples, seed):
    from oracle import synth
    secs = n_samples / synth.FS
    wav, _ = synth.fm_meeting(secs * (batch + 1), 3, seed=seed)
    ch = [[i * secs * 0.5, i * secs * 0.5 + secs] for i in range(batch)]
    return synth.cut_windows(wav, ch)


ERES_GAIN = 1.0     # weight gain of the ERes2NetV2 test networks: with He gain sqrt(2) these random
# residual stacks are chaotic (the reference's own fp32 run is 1e-1 from its fp64 run for w24s4ep4);
# at gain 1 the fp32-vs-fp64 gap is 1e-6..1e-5 while 5-10 % of layer-4 still saturates the clamp at 20


def eres2netv2_cases():
    """(name, ctor kwargs, batch, n_samples, weight seed)"""
    return [
        ("w26s2e2_t148", dict(baseWidth=26, scale=2, expansion=2), 2, 24000, 201),
        ("w24s4e4_t148", dict(baseWidth=24, scale=4, expansion=4), 2, 24000, 202),
        ("w26s2e2_t298", dict(baseWidth=26, scale=2, expansion=2), 1, 48000, 203),
    ]


ECAPA_GAIN = 1.0    # same reasoning as ERES_GAIN: keeps the random residual stack out of the chaotic regime


def import_ecapa():
    from speakerlab.models.ecapa_tdnn.ECAPA_TDNN import ECAPA_TDNN
    return ECAPA_TDNN


def ecapa_cases():
    """(name, ctor kwargs, batch, n_samples, weight seed)"""
    return [
        ("c512_t148", dict(channels=[512, 512, 512, 512, 1536]), 2, 24000, 301),
        ("c1024_t148", dict(channels=[1024, 1024, 1024, 1024, 3072]), 2, 24000, 302),
        ("c512_t298", dict(channels=[512, 512, 512, 512, 1536]), 1, 48000, 303),
    ]


def headline_cases():
    """The shapes BASELINE.json quotes its configs on, kept out of the files above so those stay byte-stable:
    (family, name, ctor kwargs, batch, n_samples, weight seed).  Config 3: ERes2NetV2 w24s4ep4 on 3 s segments
    (T=298); config 5: ECAPA-TDNN C=1024 on 10 s chunks (T=998)."""
    return [
        ("eres2netv2", "w24s4e4_t298", dict(baseWidth=24, scale=4, expansion=4), 1, 48000, 204),
        ("ecapa", "c1024_t998", dict(channels=[1024, 1024, 1024, 1024, 3072]), 1, 160000, 304),
    ]


def eres2net_cases():
    """ERes2Net v1: (name, variant, ctor kwargs of the b200spk mirror, batch, n_samples, weight seed).
    base = speakerlab.models.eres2net.ERes2Net.ERes2Net(); large = the same with m_channels=64;
    huge = speakerlab.models.eres2net.ERes2Net_huge.ERes2Net()."""
    return [
        ("base_t148", "base", dict(), 2, 24000, 401),
        ("base_t298", "base", dict(), 1, 48000, 402),
        ("large_t148", "large", dict(m_channels=64), 1, 24000, 403),
        ("huge_t148", "huge", dict(m_channels=64, baseWidth=24, scale=3, expansion=4), 1, 24000, 404),
    ]


def cluster_cases():
    """(name, N, D, K, seed, ctor kwargs)"""
    return [
        ("n600_k4", 600, 192, 4, 21, dict(min_num_spks=1, max_num_spks=15, pval=0.012)),
        ("n300_k3_default", 300, 64, 3, 22, dict()),
        ("n1500_k6", 1500, 192, 6, 23, dict(min_num_spks=1, max_num_spks=15, pval=0.012)),
    ]


def common_cases():
    """(name, N, D, K, seed, noise, outliers, CommonClustering kwargs)"""
    return [
        ("ahc_n30_k3", 30, 64, 3, 31, 0.35, 0, dict(cluster_type="AHC", fix_cos_thr=0.4)),
        ("ahc_n500_k5", 500, 192, 5, 32, 0.35, 0, dict(cluster_type="AHC", fix_cos_thr=0.4)),
        ("spectral_short_n25", 25, 64, 2, 33, 0.35, 0, dict(cluster_type="spectral", min_num_spks=1, max_num_spks=15, pval=0.012)),
        ("spectral_n400_minor_merge", 400, 96, 4, 34, 0.35, 3, dict(cluster_type="spectral", mer_cos=0.8, min_cluster_size=4,
                                                                    min_num_spks=1, max_num_spks=15, pval=0.012)),
        ("ahc_n300_minor", 300, 96, 4, 35, 0.5, 5, dict(cluster_type="AHC", fix_cos_thr=0.3, min_cluster_size=4, mer_cos=0.9)),
    ]


def common_input(n, d, k, seed, noise, outliers):
    """Gaussian blobs plus a few isolated points (they form minor clusters that filter_minor_cluster reassigns)."""
    rng = np.random.default_rng([seed, 0xC2])
    centers = rng.standard_normal((k, d))
    lab = rng.integers(0, k, n)
    X = centers[lab] + noise * rng.standard_normal((n, d))
    for j in range(outliers):
        X[j] = 2.0 * rng.standard_normal(d)
    return X.astype(np.float32), lab


def cluster_input(n, d, k, seed):
    rng = np.random.default_rng([seed, 0xC1])
    centers = rng.standard_normal((k, d))
    lab = rng.integers(0, k, n)
    X = centers[lab] + 0.35 * rng.standard_normal((n, d))
    return X.astype(np.float32), lab


def main():
    import torch
    from oracle import synth
    os.makedirs(OUT, exist_ok=True)
    FBank, CAMPPlus, SpectralCluster = import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ver = versions()

    # ---- fbank: reference fp32 and the same code on float64 input
    fb = FBank(80, 16000, mean_nor=True)
    fb_raw = FBank(80, 16000, mean_nor=False)
    out = {"versions": ver}
    for name, wavs in fbank_cases().items():
        f32 = np.stack([fb(torch.from_numpy(w).unsqueeze(0)).numpy() for w in wavs])
        f64 = np.stack([fb(torch.from_numpy(w).double().unsqueeze(0)).numpy() for w in wavs])
        raw = np.stack([fb_raw(torch.from_numpy(w).unsqueeze(0)).numpy() for w in wavs])
        # vmap form used by the diarization call sites must agree with the loop
        vm = torch.vmap(fb)(torch.from_numpy(wavs).unsqueeze(1)).numpy()
        assert np.array_equal(vm, f32), name
        out[name + ".f32"] = f32
        out[name + ".f64"] = f64
        out[name + ".raw_f32"] = raw
    np.savez_compressed(os.path.join(OUT, "fbank.npz"), **out)
    print("fbank goldens:", {k: v.shape for k, v in out.items() if k != "versions"})

    # ---- CAM++: state_dict layout + embeddings with seeded weights
    layouts = {}
    out = {"versions": ver}
    for name, emb, batch, n_samples, wseed, bnrand in campplus_cases():
        torch.manual_seed(0)
        model = CAMPPlus(feat_dim=80, embedding_size=emb).eval()
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        layouts["campplus_e%d" % emb] = {k: list(v) for k, v in shapes.items()}
        sd = synth.fill_state_dict(shapes, wseed, randomize_bn=bnrand)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        wavs = campplus_input(batch, n_samples, seed=wseed + 1000)
        feats = torch.vmap(fb)(torch.from_numpy(wavs).unsqueeze(1))
        taps = {}
        hooks = []
        for mod_name in ("head", "xvector.tdnn", "xvector.block1", "xvector.transit1", "xvector.block2",
                         "xvector.transit2", "xvector.block3", "xvector.transit3", "xvector.stats"):
            mod = model.get_submodule(mod_name)
            hooks.append(mod.register_forward_hook(
                lambda m, i, o, n=mod_name: taps.__setitem__(n, o.detach().clone())))
        with torch.no_grad():
            e = model(feats)
        for h in hooks:
            h.remove()
        out[name + ".feats"] = feats.numpy()
        out[name + ".emb"] = e.numpy()
        for k, v in taps.items():
            # per-tap fingerprint: mean, mean|x|, L2 norm  (full tensors would be MBs)
            out[name + ".tap." + k] = np.array([v.mean().item(), v.abs().mean().item(),
                                                v.double().norm().item()], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "campplus.npz"), **out)
    with open(os.path.join(OUT, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f)
    print("campplus goldens:", [k for k in out if k.endswith(".emb")])

    # ---- ERes2NetV2
    ERes2NetV2 = import_eres2netv2()
    out = {"versions": ver}
    for name, kw, batch, n_samples, wseed in eres2netv2_cases():
        torch.manual_seed(0)
        model = ERes2NetV2(feat_dim=80, embedding_size=192, **kw).eval()
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        layouts["eres2netv2_w%ds%de%d" % (kw["baseWidth"], kw["scale"], kw["expansion"])] = {k: list(v) for k, v in shapes.items()}
        sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=ERES_GAIN)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        wavs = campplus_input(batch, n_samples, seed=wseed + 1000)
        feats = torch.vmap(fb)(torch.from_numpy(wavs).unsqueeze(1))
        with torch.no_grad():
            e = model(feats.clone())
        out[name + ".feats"] = feats.numpy()
        out[name + ".emb"] = e.numpy()
    np.savez_compressed(os.path.join(OUT, "eres2netv2.npz"), **out)
    with open(os.path.join(OUT, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f)
    print("eres2netv2 goldens:", [k for k in out if k.endswith(".emb")])

    # ---- ECAPA-TDNN
    ECAPA = import_ecapa()
    out = {"versions": ver}
    for name, kw, batch, n_samples, wseed in ecapa_cases():
        torch.manual_seed(0)
        model = ECAPA(80, lin_neurons=192, **kw).eval()
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        layouts["ecapa_c%d" % kw["channels"][0]] = {k: list(v) for k, v in shapes.items()}
        sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=ECAPA_GAIN)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        wavs = campplus_input(batch, n_samples, seed=wseed + 1000)
        feats = torch.vmap(fb)(torch.from_numpy(wavs).unsqueeze(1))
        with torch.no_grad():
            e = model(feats)
            e64 = model.double()(feats.double())
        out[name + ".feats"] = feats.numpy()
        out[name + ".emb"] = e.numpy()
        out[name + ".emb_f64"] = e64.numpy()
    np.savez_compressed(os.path.join(OUT, "ecapa.npz"), **out)
    with open(os.path.join(OUT, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f)
    print("ecapa goldens:", [k for k in out if k.endswith(".emb")])

    # ---- SpectralCluster
    out = {"versions": ver}
    for name, n, d, k, seed, kw in cluster_cases():
        X, _ = cluster_input(n, d, k, seed)
        sc = SpectralCluster(**kw)
        np.random.seed(0)
        labels = sc(X.copy())
        # stage goldens
        A = sc.get_sim_mat(X)
        P = sc.p_pruning(A.copy(), None)
        sym = 0.5 * (P + P.T)
        L = sc.get_laplacian(sym)
        np.random.seed(0)
        _, kk = sc.get_spec_embs(L)
        import scipy.sparse.linalg
        lambdas = scipy.sparse.linalg.eigsh(L, k=min(sc.max_num_spks + 1, n), which="SM")[0]
        out[name + ".labels"] = labels.astype(np.int32)
        out[name + ".k"] = np.array(kk)
        out[name + ".lambdas"] = lambdas
        out[name + ".nnz_per_row"] = (P != 0).sum(1).astype(np.int32)
        out[name + ".lap_diag"] = np.diag(L).copy()
        out[name + ".lap_fro"] = np.array(np.linalg.norm(L.astype(np.float64)))
    np.savez_compressed(os.path.join(OUT, "cluster.npz"), **out)
    print("cluster goldens:", {k: out[k].tolist() for k in out if k.endswith(".k")})
    mint_common(ver)


def mint_common(ver):
    """AHCluster / CommonClustering goldens from the imported reference; fastcluster (not installed) is replaced by
    scipy.cluster.hierarchy.linkage, which implements the same average-linkage algorithm and output format."""
    import scipy.cluster.hierarchy
    fc = sys.modules["fastcluster"]
    fc.linkage = lambda y, method="average", preserve_input=True: scipy.cluster.hierarchy.linkage(y, method=method)
    from speakerlab.process import cluster as ref_cluster
    ref_cluster.fastcluster = fc
    out = {"versions": ver}
    for name, n, d, k, seed, noise, outliers, kw in common_cases():
        X, _ = common_input(n, d, k, seed, noise, outliers)
        np.random.seed(0)
        labels = ref_cluster.CommonClustering(**kw)(X.copy())
        out[name + ".labels"] = np.asarray(labels).astype(np.int32)
        if kw["cluster_type"] == "AHC":
            out[name + ".raw_ahc"] = np.asarray(ref_cluster.AHCluster(kw.get("fix_cos_thr", 0.4))(X.copy())).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "common_clustering.npz"), **out)
    print("common clustering goldens:", {k: int(v.max()) + 1 for k, v in out.items() if k.endswith(".labels")})


def mint_eres2net(ver=None):
    import torch
    from oracle import synth
    from speakerlab.models.eres2net.ERes2Net import ERes2Net as Base
    from speakerlab.models.eres2net.ERes2Net_huge import ERes2Net as Huge
    FBank, _, _ = import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    fb = FBank(80, 16000, mean_nor=True)
    out = {"versions": ver or versions()}
    layouts = json.load(open(os.path.join(OUT, "state_dict_layouts.json")))
    for name, variant, kw, batch, n_samples, wseed in eres2net_cases():
        torch.manual_seed(0)
        if variant == "huge":
            model = Huge(feat_dim=80, embedding_size=192).eval()
        else:
            model = Base(feat_dim=80, embedding_size=192, m_channels=kw.get("m_channels", 32)).eval()
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        layouts["eres2net_" + variant] = {k: list(v) for k, v in shapes.items()}
        sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=ERES_GAIN)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        wavs = campplus_input(batch, n_samples, seed=wseed + 1000)
        feats = torch.vmap(fb)(torch.from_numpy(wavs).unsqueeze(1))
        with torch.no_grad():
            e = model(feats.clone())
            e64 = model.double()(feats.double().clone())
        out[name + ".feats"] = feats.numpy()
        out[name + ".emb"] = e.numpy()
        print(name, "params %.2f M" % (sum(int(np.prod(v)) for k, v in shapes.items() if "num_batches" not in k and "running" not in k) / 1e6),
              "fp32 vs fp64 rel-L2 %.2e" % (np.linalg.norm(e.numpy() - e64.numpy()) / np.linalg.norm(e64.numpy())))
    np.savez_compressed(os.path.join(OUT, "eres2net.npz"), **out)
    with open(os.path.join(OUT, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f)


def mint_headline(ver=None):
    import torch
    from oracle import synth
    FBank, _, _ = import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    fb = FBank(80, 16000, mean_nor=True)
    out = {"versions": ver or versions()}
    for family, name, kw, batch, n_samples, wseed in headline_cases():
        torch.manual_seed(0)
        if family == "eres2netv2":
            model = import_eres2netv2()(feat_dim=80, embedding_size=192, **kw).eval()
            gain = ERES_GAIN
        else:
            model = import_ecapa()(80, lin_neurons=192, **kw).eval()
            gain = ECAPA_GAIN
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=gain)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        wavs = campplus_input(batch, n_samples, seed=wseed + 1000)
        feats = torch.vmap(fb)(torch.from_numpy(wavs).unsqueeze(1))
        with torch.no_grad():
            e = model(feats.clone())
            e64 = model.double()(feats.double().clone())
        out[name + ".feats"] = feats.numpy()
        out[name + ".emb"] = e.numpy()
        out[name + ".emb_f64"] = e64.numpy()
        print(name, "fp32 vs fp64 rel-L2 %.2e" % (np.linalg.norm(e.numpy() - e64.numpy()) / np.linalg.norm(e64.numpy())))
    np.savez_compressed(os.path.join(OUT, "headline_shapes.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "headline":
        import_reference()
        mint_headline()
    elif len(sys.argv) > 1 and sys.argv[1] == "eres2net":
        import_reference()
        mint_eres2net()
    else:
        main()
