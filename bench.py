#!/usr/bin/env python
"""Benchmark of the embedding-extraction hot path (BASELINE.json metric: CAM++ embeddings/sec on
1.5 s windows; diarization RTF) plus one measured block per remaining BASELINE config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--legs a,b,...]

Headline (the JSON line's own keys): one step = one pass of  waveform windows -> fbank+CMN -> CAM++ forward ->
embeddings  over ``--segments`` synthetic 1.5 s windows per GPU (weak scaling: every rank extracts its own slice,
no data-path collective).  ``value`` is timed with inputs resident in HBM, ``e2e`` through the public API with
pinned host buffers (int16 PCM windows, H2D of the windows and D2H of the embeddings inside the timed region).

Extra blocks under ``"configs"`` (rank 0 prints everything on ONE JSON line):
  fbank_1024x3s            BASELINE config 2: the front end alone, HBM roofline
  eres2netv2_*_4096x3s     BASELINE config 3: both ERes2NetV2 variants, bf16, device-resident and end to end, CPU subsample
  campplus_batch64         the reference call sites' own batch size (infer_diarization.py:629-635)
  campplus_fp32_mode       the same network in the fp32 precision mode (3xTF32 tensor-core convs, csrc/conv_f32x3.cu):
                           embeddings/s and how far the bf16 embeddings are from it
  ecapa_bulk_100h          BASELINE config 5: 36,000 x 10 s chunks sharded over the ranks, per-recording mean,
                           1 M trial pairs scored
  diarization              BASELINE config 4: a synthetic 1-hour meeting end to end, stage breakdown, oracle check of
                           the labels at every world size, CPU pipeline baseline
``--impl reference`` times the CPU oracle port of the headline path (torch fp32 on all host cores) on a bounded
sample per step.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "3d-speaker_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# stdout carries exactly ONE JSON line.  Native libraries write there too (NCCL prints "NCCL version ..." with printf
# when the box sets NCCL_DEBUG): keep a private handle on the real stdout for the result line and point file descriptor
# 1 at stderr for everything else.
_RESULT_OUT = None


def claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


N_SAMPLES = 24000                 # 1.5 s @ 16 kHz  (infer_diarization.py:285 chunk_dur)
T_FRAMES = 148
EMB = 512                         # CAM++ 7.2 M variant (BASELINE config 0)
GFLOP_PER_SEG = 1.588             # minimal conv/linear FLOPs per 1.5 s segment
FBANK_BYTES_PER_SEG = 4 * N_SAMPLES + 4 * T_FRAMES * 80
WEIGHT_SEED = 7
ALL_LEGS = ("fbank", "eres", "batch64", "fp32", "ecapa", "meeting")
METRIC = "CAM++ embeddings/sec (1.5 s windows)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def source_hash():
    """sha256 over the CUDA sources: ties a profile under profiles/ to the code that produced it."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "3d-speaker_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes per segment of one CAM++ forward call, written by tools/measure_traffic.py from an ncu pass
    (dram__bytes_read.sum + dram__bytes_write.sum over every kernel of the call).  None when no profile of the
    CURRENT sources is committed."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu traffic profile committed"
    t = json.load(open(path))
    note = "ncu dram__bytes over one forward call, profiles/r02_traffic.json (tools/measure_traffic.py)"
    if t.get("source_hash") != source_hash():
        note += "; measured on an earlier build of csrc/ (source hash differs)"
    return t["dram_bytes_per_segment"], note


def make_weights(model, gain=None):
    from oracle import synth
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    kw = {} if gain is None else {"gain": gain}
    sd = synth.fill_state_dict(shapes, WEIGHT_SEED, randomize_bn=True, **kw)
    return {k: torch.from_numpy(v) for k, v in sd.items()}, sd


def make_windows(n, seed, n_samples=N_SAMPLES):
    """Seeded synthetic 16 kHz audio cut into windows: noise bursts with a noise floor (cheap to generate for
    thousands of windows; the parity tests use the FM-speaker meeting).  float32 in [-1, 1] scale."""
    rng = np.random.default_rng([seed, 0xBE7C])
    x = rng.standard_normal((n, n_samples), dtype=np.float32)
    env = 0.05 + 0.15 * np.abs(np.sin(np.linspace(0, 9.0, n_samples, dtype=np.float32)))[None, :]
    return (x * env).astype(np.float32)


def to_pcm(x):
    """float [-1, 1] -> int16 PCM, what a 16-bit wav file holds (and what fileio.py:115-117 divides by 32768)."""
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


class ClockSampler:
    """SM clock + throttle-reason sampler running DURING the timed region (B200_PROFILING.md clocks
    line).  NVML in-process (pynvml) every 20 ms; falls back to polling nvidia-smi."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.sm, self.max_sm, self.reasons = index, [], None, set()
        self._stop, self._t, self.src = threading.Event(), None, "nvml"

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        try:                          # map the torch device index to NVML through the UUID (CUDA_VISIBLE_DEVICES may remap)
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self._stop.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
            self._stop.wait(0.02)

    def _run_smi(self):
        self.src = "nvidia-smi"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    c = [x.strip() for x in out.stdout.strip().split(",")]
                    self.sm.append(float(c[0]))
                    self.max_sm = float(c[1])
                    self.reasons.update(n for i, n in enumerate(names) if c[2 + i].lower().startswith("active"))
            except Exception:
                pass
            self._stop.wait(0.1)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0, "source": self.src}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": self.src}


def ev():
    return torch.cuda.Event(enable_timing=True)


class Ctx:
    """Process-wide state of the GPU arm."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.pk = peaks()
        self.cores = os.cpu_count() or 1

    def barrier(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = torch.tensor([float(v)], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def workload_config(args):
    return {"workload": "CAM++ (7.2M, emb 512, random-init) embedding extraction: fbank+CMN -> forward on synthetic "
                        "16 kHz 1.5 s windows (BASELINE config 0 shape)",
            "segments_per_gpu_per_step": args.segments,
            "parallelism": "dp%d (sub-segment sharding, no collective)" % args.gpus,
            "l2": "inputs larger than L2 (%.0f MB of float32 windows per step)" % (args.segments * N_SAMPLES * 4 / 1e6)}


# ============================================================================= CPU oracle legs
def cpu_campplus(seconds_budget, batch=64):
    """The oracle port of the headline path (numpy fbank + torch fp32 CAM++) on all host cores, on a bounded
    sample of the same workload, at the reference's own batch size."""
    from oracle import campplus_oracle, fbank_oracle
    import b200spk
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = b200spk.CAMPPlus(embedding_size=EMB)          # parameter shapes only (never run on CPU)
    _, sd = make_weights(model)
    wavs = make_windows(batch, seed=99)
    campplus_oracle.forward(sd, fbank_oracle.fbank_batch(wavs[:2]))
    done, t0 = 0, time.perf_counter()
    while True:
        campplus_oracle.forward(sd, fbank_oracle.fbank_batch(wavs))
        done += batch
        el = time.perf_counter() - t0
        if el >= seconds_budget:
            break
    return dict(value=done / el, unit="embeddings/s", cores=cores, kind="port",
                sample="%d x 1.5 s windows in batches of %d, fp32 (%.1f s of CPU work)" % (done, batch, el))


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    per_step = 10.0
    from oracle import campplus_oracle, fbank_oracle
    import b200spk
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = b200spk.CAMPPlus(embedding_size=EMB)
    _, sd = make_weights(model)
    batch = 64
    wavs = make_windows(batch, seed=99)

    def step():
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < per_step:
            campplus_oracle.forward(sd, fbank_oracle.fbank_batch(wavs))
            n += batch
        return n, time.perf_counter() - t0
    for _ in range(min(args.warmup, 1)):
        step()
    steps_run = max(1, min(args.steps, 12))
    tot_n, tot_t = 0, 0.0
    for _ in range(steps_run):
        n, t = step()
        tot_n, tot_t = tot_n + n, tot_t + t
    val = tot_n / tot_t
    sample = ("%d x 1.5 s windows per ~%.0f s step (%d steps run), batches of %d, fp32, numpy fbank + torch CAM++ "
              "(oracle port)" % (tot_n // steps_run, per_step, steps_run, batch))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "embeddings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_run": steps_run,
        "ms_per_step": 1e3 * tot_t / steps_run,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),          # the workload; what this arm ran of it is in cpu_baseline.sample
        "arm": {"batch": batch, "precision": "fp32", "device": "host cores"},
        "cpu_baseline": {"value": val, "unit": "embeddings/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ============================================================================= config 2: fbank alone
def leg_fbank(cx):
    """1024 x 3 s of 0.1 * randn (SURVEY 8d set A), HBM-resident; L2 flushed between iterations; parity numbers on a
    64-utterance subsample against the float64 oracle."""
    import b200spk
    from oracle import fbank_oracle
    torch.manual_seed(1)
    wav = (0.1 * torch.randn(1024, 48000)).to(cx.dev)
    fb = b200spk.FBank(80, 16000, mean_nor=True)
    for _ in range(5):
        out = fb.batch(wav)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=cx.dev)     # 256 MB > L2
    ms = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = ev(), ev()
        e0.record()
        out = fb.batch(wav)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms))
    nbytes = 1024 * (4 * 48000 + 4 * 298 * 80)
    sub = wav[:64].cpu().numpy()
    t0 = time.perf_counter()
    r64 = fbank_oracle.fbank_batch(sub, dtype=np.float64)
    cpu_s = time.perf_counter() - t0
    err = np.abs(out[:64].cpu().numpy() - r64)
    gbs = nbytes / t / 1e6
    return {"workload": "Kaldi 80-dim fbank + CMN alone, 1024 x 3 s utterances (0.1 * randn), HBM-resident, L2 flushed",
            "ms": t, "utterances_per_s": 1024 / t * 1e3, "gpu_launches_per_call": 1,
            "roofline": {"kernel": "fbank_kernel", "bound": "hbm", "achieved": gbs, "peak": cx.pk["hbm"], "unit": "GB/s",
                         "frac": gbs / cx.pk["hbm"], "bytes_per_launch": nbytes, "traffic": None,
                         "note": "traffic = algorithmic minimum (one read of the samples, one write of the features); "
                                 "the kernel is fp32-pipe bound, see profiles/r02_ncu_fbank.md"},
            "parity": {"max_abs_vs_float64": float(err.max()), "mean_abs_vs_float64": float(err.mean()), "utterances_checked": 64},
            "cpu_baseline": {"value": 64 / cpu_s, "unit": "utterances/s", "cores": 1, "kind": "port",
                             "sample": "64 x 3 s utterances, numpy float64 oracle, per-utterance loop"}}


# ============================================================================= config 3: ERes2NetV2
def leg_eres(cx):
    import b200spk
    from oracle import eres2netv2_oracle
    out = {}
    n_seg, T, n_samples = cx.args.eres_segments, 298, 48000
    wav_f = make_windows(n_seg, seed=31, n_samples=n_samples)
    host = torch.from_numpy(to_pcm(wav_f)).pin_memory()
    wav_dev = torch.from_numpy(wav_f).to(cx.dev)
    del wav_f
    fb = b200spk.FBank(80, 16000, mean_nor=True)
    for name, kw, gflop, batch in (("eres2netv2_w24s4ep4_4096x3s", dict(baseWidth=24, scale=4, expansion=4), 73.850, 128),
                                   ("eres2netv2_w26s2e2_4096x3s", dict(baseWidth=26, scale=2, expansion=2), 24.934, 256)):
        model = b200spk.ERes2NetV2(precision="bf16", **kw)
        tsd, sd = make_weights(model, gain=1.0)
        model.load_state_dict(tsd)
        model = model.to(cx.dev).eval()
        ex = b200spk.EmbeddingExtractor(fb, model, device=cx.dev, batchsize=batch, reuse_output=True, head=batch)
        L = b200spk.lib()
        with torch.no_grad():
            ex.extract_device(wav_dev[:2 * batch])
            torch.cuda.synchronize()
            l0 = L.spk_launch_count()
            e0, e1 = ev(), ev()
            e0.record()
            emb = ex.extract_device(wav_dev)
            e1.record()
            torch.cuda.synchronize()
            launches = L.spk_launch_count() - l0
        ms = e0.elapsed_time(e1)
        ex(host[:2 * batch])
        t0 = time.perf_counter()
        emb_host = ex(host)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert emb_host.shape == (n_seg, 192) and bool(torch.isfinite(emb_host).all())
        # CPU subsample: the fp32 oracle port of the network on the same features (also the parity check)
        k = 8
        feats = fb.batch(wav_dev[:k]).cpu().numpy()
        torch.set_num_threads(cx.cores)
        eres2netv2_oracle.forward(sd, feats[:1], scale=kw["scale"])
        t0 = time.perf_counter()
        ref = eres2netv2_oracle.forward(sd, feats, scale=kw["scale"]).numpy()
        cpu_s = time.perf_counter() - t0
        got = emb[:k].cpu().numpy()
        cos = float(((got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))).min())
        tf = n_seg * gflop / ms
        out[name] = {"workload": "ERes2NetV2 baseWidth=%d scale=%d expansion=%d (random-init, randomised BN), %d x 3 s segments, bf16, "
                                 "fbank+CMN -> forward" % (kw["baseWidth"], kw["scale"], kw["expansion"], n_seg),
                     "segments_per_s": n_seg / ms * 1e3, "ms": ms, "batch": batch, "gpu_launches": int(launches),
                     "e2e": {"value": n_seg / e2e_s, "unit": "segments/s", "h2d_bytes": n_seg * n_samples * 2,
                             "d2h_bytes": n_seg * 192 * 4, "input": "int16 PCM in pinned host memory"},
                     "roofline": {"kernel": "conv stack of one extraction pass (fbank included in the time)", "bound": "tensor",
                                  "achieved": tf, "peak": cx.pk["bf16_sustained"], "unit": "TFLOP/s",
                                  "frac": tf / cx.pk["bf16_sustained"], "gflop_per_segment": gflop, "traffic": None},
                     "cpu_baseline": {"value": k / cpu_s, "unit": "segments/s", "cores": cx.cores, "kind": "port",
                                      "sample": "%d x 3 s segments, torch fp32 oracle port of the forward (features from the GPU)" % k},
                     "parity": {"min_cos_bf16_vs_cpu_fp32": cos, "segments_checked": k,
                                "note": "seeded weights with randomised BN statistics (stress set); random-init is >= 0.9999, tests/test_gpu_bf16_parity.py"}}
        del model, ex
        torch.cuda.empty_cache()
    return out


# ============================================================================= batch 64
def leg_batch64(cx, model, fb):
    """The reference call sites batch 64 windows per fbank/forward call (infer_diarization.py:629-635): the same
    extractor at that batch size, host buffers in and out."""
    import b200spk
    n = 4096
    host = torch.from_numpy(to_pcm(make_windows(n, seed=3))).pin_memory()
    out = {}
    for bs in (64, 256):
        ex = b200spk.EmbeddingExtractor(fb, model, device=cx.dev, batchsize=bs, reuse_output=True, head=bs)
        for _ in range(2):
            ex(host)
        torch.cuda.synchronize()
        l0 = b200spk.lib().spk_launch_count()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ex(host)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        calls = (n + bs - 1) // bs
        out["batch%d" % bs] = {"embeddings_per_s_e2e": n / dt, "ms_per_call": dt / calls * 1e3,
                               "gpu_launches_per_call": int((b200spk.lib().spk_launch_count() - l0) // reps // calls)}
    out["workload"] = "CAM++ 512-d bf16, %d x 1.5 s int16 windows through EmbeddingExtractor at the reference's batch size" % n
    return out


# ============================================================================= fp32 precision mode
def leg_fp32(cx, model_bf16, fb):
    """precision="fp32": every conv as a 3xTF32 split product on the tensor cores with fp32 register sums
    (csrc/conv_f32x3.cu) - the mode whose embeddings match the reference's CPU fp32 forward to 1e-4 rel-L2 (tests/)."""
    import b200spk
    n, bs = 2048, 1024
    m32 = b200spk.CAMPPlus(embedding_size=EMB, precision="fp32")
    m32.load_state_dict(model_bf16.state_dict())
    m32 = m32.to(cx.dev).eval()
    wav = torch.from_numpy(make_windows(n, seed=4)).to(cx.dev)
    with torch.no_grad():
        feats = fb.batch(wav)

        def run(m):
            return torch.cat([m(feats[i:i + bs]) for i in range(0, n, bs)])
        for _ in range(2):
            e32 = run(m32)
        torch.cuda.synchronize()
        l0 = b200spk.lib().spk_launch_count()
        t0, t1 = ev(), ev()
        t0.record()
        reps = 3
        for _ in range(reps):
            e32 = run(m32)
        t1.record()
        torch.cuda.synchronize()
        dt = t0.elapsed_time(t1) / reps * 1e-3
        launches = int((b200spk.lib().spk_launch_count() - l0) // reps)
        e16 = run(model_bf16)
    cos = torch.nn.functional.cosine_similarity(e16.float(), e32.float(), dim=1)
    out = {"embeddings_per_s": n / dt, "ms_per_step": dt * 1e3, "gpu_launches": launches, "dtype": "f32 (3xTF32 products, fp32 sums)",
           "tflops_fp32_equivalent": GFLOP_PER_SEG * n / dt / 1e3,
           "bf16_vs_fp32_mode": {"cos_min": float(cos.min()), "rel_l2": float((e16 - e32).norm() / e32.norm())},
           "workload": "CAM++ 512-d, %d x 1.5 s segments, feature matrices resident in HBM, forward calls of %d" % (n, bs)}
    del m32
    torch.cuda.empty_cache()
    return out


# ============================================================================= config 5: bulk ECAPA + scoring
def leg_ecapa(cx):
    """100 h of synthetic audio = 4,000 recordings of 90 s -> 36,000 x 10 s chunks (infer_sv_batch.py:388-412),
    recordings sharded over the ranks, ECAPA-TDNN C=1024 bf16, per-recording mean, all_gather, 1 M trial pairs."""
    import b200spk
    n_wav_total = cx.args.bulk_recordings
    assert n_wav_total % cx.world == 0, "--bulk-recordings must divide by the number of ranks"
    lo, hi = b200spk.shard_range(n_wav_total, cx.rank, cx.world)
    n_wav = hi - lo
    rec_len = 90 * 16000
    model = b200spk.ECAPA_TDNN(80, lin_neurons=192, channels=[1024, 1024, 1024, 1024, 3072], precision="bf16")
    tsd, sd = make_weights(model, gain=1.0)
    model.load_state_dict(tsd)
    model = model.to(cx.dev).eval()
    fb = b200spk.FBank(80, 16000, mean_nor=True)
    bx = b200spk.BulkExtractor(fb, model, device=cx.dev, batchsize=128)
    g = torch.Generator(device=cx.dev).manual_seed(500 + cx.rank)
    buf = torch.empty(n_wav * rec_len, dtype=torch.int16, device=cx.dev)          # the shard's audio, resident as PCM
    slab = 50 * rec_len
    for o in range(0, buf.numel(), slab):
        n = min(slab, buf.numel() - o)
        buf[o:o + n] = (torch.randn(n, generator=g, device=cx.dev) * 3000.0).clamp_(-32768, 32767).to(torch.int16)
    lengths = [rec_len] * n_wav
    n_chunks = n_wav * 9
    L = b200spk.lib()
    bx(buf[:8 * rec_len], lengths[:8])                                             # warm-up: compile T=998
    cx.barrier()
    l0 = L.spk_launch_count()
    e0, e1 = ev(), ev()
    e0.record()
    chunk_emb, pos = bx.chunk_embeddings(buf, lengths)
    wav_emb = b200spk.segment_mean(chunk_emb, pos)
    e1.record()
    cx.barrier()
    launches = L.spk_launch_count() - l0
    ms = cx.max_over_ranks(e0.elapsed_time(e1))
    # end to end: recordings stream from pinned host memory in slabs of 100, embeddings come back per recording
    per = min(100, n_wav)
    hslab = [torch.empty(per * rec_len, dtype=torch.int16).pin_memory() for _ in range(2)]
    hslab[0].copy_(buf[:per * rec_len].cpu())
    hslab[1].copy_(hslab[0])
    copy_stream = torch.cuda.Stream(cx.dev)
    main = torch.cuda.current_stream(cx.dev)
    dslab = [torch.empty(per * rec_len, dtype=torch.int16, device=cx.dev) for _ in range(2)]
    out_host = torch.empty((n_wav, 192), dtype=torch.float32).pin_memory()
    cx.barrier()
    t0 = time.perf_counter()
    evs = [None, None]
    done = [None, None]

    def stage(i, k):
        with torch.cuda.stream(copy_stream):
            if done[k] is not None:
                copy_stream.wait_event(done[k])
            dslab[k].copy_(hslab[k], non_blocking=True)
            evs[k] = torch.cuda.Event()
            evs[k].record(copy_stream)
    n_slabs = (n_wav + per - 1) // per
    stage(0, 0)
    for i in range(n_slabs):
        k = i & 1
        if i + 1 < n_slabs:
            stage(i + 1, (i + 1) & 1)
        main.wait_event(evs[k])
        cnt = min(per, n_wav - i * per)
        e = bx(dslab[k][:cnt * rec_len], lengths[:cnt])
        done[k] = torch.cuda.Event()
        done[k].record(main)
        out_host[i * per:i * per + cnt].copy_(e, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = cx.max_over_ranks(time.perf_counter() - t0)
    # gather the chunk embeddings for trial scoring (the only collective of this config)
    all_emb = b200spk.gather_embeddings(chunk_emb, n_wav_total * 9)
    res = None
    n_pairs = 1_000_000
    rng = np.random.default_rng(5)
    a = torch.from_numpy(rng.integers(0, n_wav_total * 9, n_pairs, dtype=np.int32)).to(cx.dev)
    b = torch.from_numpy(rng.integers(0, n_wav_total * 9, n_pairs, dtype=np.int32)).to(cx.dev)
    for _ in range(3):
        sc = b200spk.cosine_pairs(all_emb, a, b)
    sms = []
    for _ in range(10):
        s0, s1 = ev(), ev()
        s0.record()
        sc = b200spk.cosine_pairs(all_emb, a, b)
        s1.record()
        torch.cuda.synchronize()
        sms.append(s0.elapsed_time(s1))
    if cx.rank == 0:
        score_ms = float(np.median(sms))
        gflop = 35.848
        tf = n_wav_total * 9 * gflop / ms
        res = {"workload": "ECAPA-TDNN C=1024 (20.8M, random-init, randomised BN) bulk extraction: %d recordings x 90 s = %d x 10 s chunks "
                           "(%.0f h), bf16, recordings sharded over %d rank(s); per-recording mean; 1 M trial pairs"
                           % (n_wav_total, n_wav_total * 9, n_wav_total * 90 / 3600.0, cx.world),
               "chunks_per_s": n_wav_total * 9 / ms * 1e3, "audio_hours_per_s": n_wav_total * 90 / 3600.0 / (ms / 1e3), "ms": ms,
               "gpu_launches": int(launches), "resident_audio_bytes_per_gpu": int(buf.numel() * 2),
               "e2e": {"value": n_wav_total * 9 / e2e_s, "unit": "chunks/s", "h2d_bytes": n_wav_total * rec_len * 2,
                       "d2h_bytes": n_wav_total * 192 * 4, "input": "int16 PCM from pinned host memory in slabs of 100 recordings"},
               "roofline": {"kernel": "conv stack (fbank and pooling included in the time)", "bound": "tensor", "achieved": tf,
                            "peak": cx.pk["bf16_sustained"] * cx.world, "unit": "TFLOP/s", "frac": tf / (cx.pk["bf16_sustained"] * cx.world),
                            "gflop_per_chunk": gflop, "traffic": None},
               "scoring": {"pairs": n_pairs, "ms": score_ms, "pairs_per_s": n_pairs / score_ms * 1e3,
                           "roofline": {"kernel": "cosine_pairs_kernel", "bound": "hbm", "achieved": n_pairs * 2 * 192 * 4 / score_ms / 1e6,
                                        "peak": cx.pk["hbm"], "unit": "GB/s", "frac": n_pairs * 2 * 192 * 4 / score_ms / 1e6 / cx.pk["hbm"],
                                        "note": "gather traffic of 2 x 192 floats per pair; the 27.6 MB embedding table itself stays in L2"}}}
        if cx.world == 1 and not cx.args.no_cpu:
            from oracle import ecapa_oracle
            from sklearn.metrics.pairwise import cosine_similarity
            k = 4
            starts, periods, phases, _ = b200spk.chunk_table(lengths[:1])
            feats = b200spk.fbank_windows(buf, torch.from_numpy(starts[:k]).to(cx.dev), torch.from_numpy(periods[:k]).to(cx.dev),
                                          160000, 80, True, torch.from_numpy(phases[:k]).to(cx.dev)).cpu().numpy()
            torch.set_num_threads(cx.cores)
            ecapa_oracle.forward(tsd, feats[:1])
            t0 = time.perf_counter()
            ref = ecapa_oracle.forward(tsd, feats).numpy()
            cpu_s = time.perf_counter() - t0
            got = chunk_emb[:k].cpu().numpy()
            cos = float(((got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))).min())
            E = all_emb.cpu().numpy()
            ah, bh = a[:10000].cpu().numpy(), b[:10000].cpu().numpy()
            t0 = time.perf_counter()
            ref_sc = np.array([cosine_similarity(E[i][None], E[j][None])[0, 0] for i, j in zip(ah, bh)])    # compute_score_metrics.py:113-114
            loop_s = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": k / cpu_s, "unit": "chunks/s", "cores": cx.cores, "kind": "port",
                                   "sample": "%d x 10 s chunks, torch fp32 oracle port of the forward" % k}
            res["scoring"]["cpu_baseline"] = {"value": 10000 / loop_s, "unit": "pairs/s", "cores": 1, "kind": "port",
                                              "sample": "10,000 pairs, sklearn cosine_similarity per trial in a Python loop"}
            res["parity"] = {"min_cos_bf16_vs_cpu_fp32": cos, "chunks_checked": k,
                             "scores_max_abs_vs_sklearn": float(np.abs(sc[:10000].cpu().numpy() - ref_sc).max())}
    del buf, chunk_emb, all_emb
    torch.cuda.empty_cache()
    return res


# ============================================================================= config 4: diarization
def leg_meeting(cx):
    """A synthetic 1-hour meeting -> 4799 sub-segments -> CAM++ (192-d) -> cosine affinity + spectral clustering,
    sub-segments sharded over the ranks, one all_gather of the embeddings.  Timed end to end from the int16 PCM
    recording in pinned host memory to labels on the host; RTF = wall / audio seconds."""
    import b200spk
    from oracle import campplus_oracle, cluster_oracle, fbank_oracle, synth
    args = cx.args
    secs, K = args.meeting_seconds, args.meeting_speakers
    n = int(round(secs * 16000))
    if cx.rank == 0:
        wav_np, turns = synth.fm_meeting(secs, K, seed=18)
        wav = torch.from_numpy(to_pcm(wav_np)).pin_memory()
    else:
        wav, turns, wav_np = torch.empty(n, dtype=torch.int16).pin_memory(), None, None
    if cx.dist is not None:
        tmp = wav.to(cx.dev)
        cx.dist.broadcast(tmp.view(torch.uint8), src=0)     # NCCL has no int16: ship the PCM as bytes
        wav.copy_(tmp)
        del tmp
    torch.manual_seed(1)                                   # default init, BN not randomised
    model = b200spk.CAMPPlus(embedding_size=192, precision=args.precision).to(cx.dev).eval()
    dz = b200spk.Diarizer(b200spk.FBank(80, 16000, mean_nor=True), model,
                          b200spk.SpectralCluster(min_num_spks=1, max_num_spks=15, pval=0.012, device=cx.dev),
                          device=cx.dev, batchsize=args.batch)
    times, stages = [], []
    for it in range(5):
        cx.barrier()
        np.random.seed(0)
        t0 = time.perf_counter()
        chunks, labels = dz(wav)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        stages.append(dict(dz.last))
    order = np.argsort(times[1:])
    med = 1 + int(order[len(order) // 2])
    wall = cx.max_over_ranks(times[med])
    if cx.rank != 0:
        return None
    truth, pure = synth.turn_labels(chunks, turns)
    purity = sum(np.bincount(truth[pure & (labels == c)]).max() for c in np.unique(labels[pure])) / pure.sum()
    # parity of the back end at full size at EVERY world size: the CPU oracle on the gathered embeddings
    emb = dz.last_embeddings.cpu().numpy()
    np.random.seed(0)
    t0 = time.perf_counter()
    ref, st = cluster_oracle.spectral_cluster(emb, 1, 15, 0.012, return_stages=True)
    cpu_cluster_s = time.perf_counter() - t0
    m = cluster_oracle.match_labels(ref, labels)
    out = {"audio_seconds": secs, "n_subsegments": len(chunks), "speakers": K, "wall_s": wall, "rtf": wall / secs,
           "input": "int16 PCM recording in pinned host memory (%.0f MB H2D inside the timed region)" % (n * 2 / 1e6),
           "k": int(dz.cluster.last["k"]), "purity_on_single_speaker_segments": float(purity),
           "krylov_dim": int(dz.cluster.last["krylov"]), "stages_s": {k: float(v) for k, v in stages[med].items()},
           "vs_oracle_backend": {"k_oracle": int(st["k"]), "mismatch_pure": int((m != ref)[pure].sum()),
                                 "mismatch_total": int((m != ref).sum()), "world_size": cx.world}}
    if cx.world == 1 and not args.no_cpu:
        # CPU baseline of the WHOLE config-4 pipeline: extraction on a bounded sample of the same windows
        # (extrapolated to 4799), clustering in full on the same embeddings (timed above)
        torch.set_num_threads(cx.cores)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        pick = chunks[:256]
        wins = synth.cut_windows(wav_np, pick)
        campplus_oracle.forward(sd, fbank_oracle.fbank_batch(wins[:2]))
        t0 = time.perf_counter()
        for i in range(0, len(pick), 64):
            campplus_oracle.forward(sd, fbank_oracle.fbank_batch(wins[i:i + 64]))
        ext_s = (time.perf_counter() - t0) * len(chunks) / len(pick)
        out["cpu_baseline"] = {"value": (ext_s + cpu_cluster_s) / secs, "unit": "RTF", "cores": cx.cores, "kind": "port",
                               "extract_s_extrapolated": ext_s, "cluster_s": cpu_cluster_s,
                               "sample": "extraction: 256 of the %d windows in batches of 64 (numpy fbank + torch fp32 CAM++ 192-d), "
                                         "extrapolated; clustering: the full oracle back end on the %d embeddings" % (len(chunks), len(chunks))}
    return out


# ============================================================================= GPU arm
def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--segments", type=int, default=16384, help="1.5 s windows per GPU per step")
    ap.add_argument("--batch", type=int, default=8192, help="windows per fbank/forward call")
    ap.add_argument("--chunk", type=int, default=0, help="coarse sub-batch (D-TDNN part), 0 = auto")
    ap.add_argument("--fine", type=int, default=0, help="fine sub-batch (2-D front), 0 = auto")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--legs", default="all", help="comma list of %s, 'all' or 'none'" % ",".join(ALL_LEGS))
    ap.add_argument("--no-meeting", action="store_true", help="skip the 1-hour diarization leg")
    ap.add_argument("--meeting-seconds", type=float, default=3600.0)
    ap.add_argument("--meeting-speakers", type=int, default=8)
    ap.add_argument("--eres-segments", type=int, default=4096)
    ap.add_argument("--bulk-recordings", type=int, default=4000, help="90 s recordings in the bulk leg (4000 = 100 h)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    legs = set(ALL_LEGS) if args.legs == "all" else set(x for x in args.legs.split(",") if x and x != "none")
    if args.no_meeting:
        legs.discard("meeting")

    import b200spk
    cx = Ctx(args)
    dev, rank, world = cx.dev, cx.rank, cx.world

    chunk = None
    if args.chunk or args.fine:
        chunk = (args.chunk or 8192, args.fine or min(args.chunk or 8192, 2048))
    model = b200spk.CAMPPlus(embedding_size=EMB, precision=args.precision, chunk=chunk)
    tsd, _ = make_weights(model)
    model.load_state_dict(tsd)
    model = model.to(dev).eval()
    fb = b200spk.FBank(80, 16000, mean_nor=True)
    ex = b200spk.EmbeddingExtractor(fb, model, device=dev, batchsize=args.batch, reuse_output=True)

    S = args.segments
    pcm = to_pcm(make_windows(S, seed=1000 + rank))
    host = torch.from_numpy(pcm).pin_memory()                                   # what a wav file holds
    wav_dev = (host.to(dev).to(torch.float32) * (1.0 / 32768.0)).contiguous()   # the same samples, float32, HBM-resident
    del pcm

    L = b200spk.lib()
    fb_ms, fw_ms = [], []

    def step_device(record):
        outs = []
        with torch.no_grad():
            for st in range(0, S, args.batch):
                wb = wav_dev[st:st + args.batch]
                if record is not None:
                    e0, e1, e2 = ev(), ev(), ev()
                    e0.record()
                feats = fb.batch(wb)
                if record is not None:
                    e1.record()
                outs.append(model(feats))
                if record is not None:
                    e2.record()
                    record.append((e0, e1, e2))
        return outs

    for _ in range(args.warmup):
        step_device(None)
    cx.barrier()
    launches0 = L.spk_launch_count()
    recs = []
    with ClockSampler(cx.local) as clk:
        t_start, t_end = ev(), ev()
        t_start.record()
        for _ in range(args.steps):
            outs = step_device(recs)
        t_end.record()
        cx.barrier()
    launches = L.spk_launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    for e0, e1, e2 in recs:
        fb_ms.append(e0.elapsed_time(e1))
        fw_ms.append(e1.elapsed_time(e2))
    ms = cx.max_over_ranks(ms)
    value = world * S * args.steps / (ms / 1e3)
    emb_dev = torch.cat(outs, dim=0)

    # ---- end to end through the public API: int16 PCM windows in pinned host memory -> embeddings on the host
    for _ in range(2):
        ex(host)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        emb_host = ex(host)
    torch.cuda.synchronize()
    e2e_s = cx.max_over_ranks(time.perf_counter() - t0)
    e2e = world * S * args.steps / e2e_s
    assert emb_host.shape == (S, EMB) and bool(torch.isfinite(emb_host).all())
    e2e_equal = bool(torch.equal(emb_host[:256], emb_dev[:256].cpu()))     # int16 host path vs float32 device path
    # the same with float32 host windows (4 bytes per sample over PCIe), for reference
    host_f = wav_dev.cpu().pin_memory()
    ex(host_f)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(max(2, args.steps // 4)):
        ex(host_f)
    torch.cuda.synchronize()
    e2e_f = world * S * max(2, args.steps // 4) / cx.max_over_ranks(time.perf_counter() - t0)
    del host_f

    configs = {}
    if world == 1:
        if "fbank" in legs:
            configs["fbank_1024x3s"] = leg_fbank(cx)
        if "batch64" in legs:
            configs["campplus_batch64"] = leg_batch64(cx, model, fb)
        if "fp32" in legs and args.precision == "bf16":
            configs["campplus_fp32_mode"] = leg_fp32(cx, model, fb)
    del wav_dev, emb_dev
    torch.cuda.empty_cache()
    if world == 1 and "eres" in legs:
        configs.update(leg_eres(cx))
    if "ecapa" in legs:
        r = leg_ecapa(cx)
        if r is not None:
            configs["ecapa_bulk_100h"] = r
    meeting = leg_meeting(cx) if "meeting" in legs else None

    if rank == 0:
        pk = cx.pk
        calls = len(fw_ms)
        fw_avg = sum(fw_ms) / calls            # ms per forward call (args.batch segments)
        fb_avg = sum(fb_ms) / calls
        segs_per_call = S / (S // args.batch + (1 if S % args.batch else 0))
        tf = GFLOP_PER_SEG * segs_per_call / fw_avg          # GFLOP/ms == TFLOP/s
        gbs = FBANK_BYTES_PER_SEG * segs_per_call / fb_avg / 1e6
        per_seg, tnote = measured_traffic() if args.precision == "bf16" else (None, "fp32 mode not profiled")
        traffic = per_seg * segs_per_call if per_seg else None
        line = {
            "metric": METRIC, "value": value, "unit": "embeddings/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args),
            "arm": {"batch": args.batch, "precision": args.precision, "device": "B200"},
            "e2e": {"value": e2e, "unit": "embeddings/s", "h2d_bytes_per_step": S * N_SAMPLES * 2,
                    "d2h_bytes_per_step": S * EMB * 4, "input": "int16 PCM windows in pinned host memory (scaled by 1/32768 in the fbank kernel)",
                    "bit_identical_to_device_path": e2e_equal},
            "e2e_float32_host": {"value": e2e_f, "unit": "embeddings/s", "h2d_bytes_per_step": S * N_SAMPLES * 4,
                                 "d2h_bytes_per_step": S * EMB * 4},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "CAM++ conv stack (all implicit-GEMM launches of one forward call)",
                         "bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": tf / pk["bf16_sustained"], "traffic": traffic, "peak_source": pk["source"] + " (sustained)",
                         "flops_per_launch": GFLOP_PER_SEG * 1e9 * segs_per_call, "ms_per_launch": fw_avg,
                         "traffic_source": tnote},
            "roofline_hbm": {"kernel": "CAM++ forward call (all launches)", "bound": "hbm",
                             "achieved": (traffic / (fw_avg * 1e-3) / 1e9) if traffic else None, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": (traffic / (fw_avg * 1e-3) / 1e9 / pk["hbm"]) if traffic else None, "traffic": traffic},
            "roofline_fbank": {"kernel": "fbank_kernel", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"],
                               "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None,
                               "bytes_per_launch": FBANK_BYTES_PER_SEG * segs_per_call, "ms_per_launch": fb_avg},
            "clocks": clk.summary(),
            "configs": configs,
        }
        if meeting is not None:
            line["diarization"] = meeting
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_campplus(args.cpu_seconds)
        emit(line)
    if cx.dist is not None:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
