#!/usr/bin/env python
"""Benchmark of the embedding-extraction hot path (BASELINE.json metric: CAM++ embeddings/sec on
1.5 s windows).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of  waveform windows -> fbank+CMN -> CAM++ forward -> embeddings  over
``--segments`` synthetic 1.5 s windows per GPU (weak scaling: every rank extracts its own slice,
no data-path collective).  Prints ONE JSON line (rank 0).  ``value`` is timed with inputs
resident in HBM, ``e2e`` through the public API with pinned host buffers (H2D of the windows and
D2H of the embeddings inside the timed region).  ``--impl reference`` times the CPU oracle port
of the same path (torch fp32 on all host cores) on a bounded sample.
"""
import argparse
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

N_SAMPLES = 24000                 # 1.5 s @ 16 kHz  (infer_diarization.py:285 chunk_dur)
T_FRAMES = 148
EMB = 512                         # CAM++ 7.2 M variant (BASELINE config 0)
DRAM_BYTES_PER_SEG_BF16 = 31.29e9 / 2048      # measured with ncu on the bf16 path (profiles/r01_traffic_o.md)
GFLOP_PER_SEG = 1.588             # minimal conv/linear FLOPs per 1.5 s segment
FBANK_BYTES_PER_SEG = 4 * N_SAMPLES + 4 * T_FRAMES * 80
WEIGHT_SEED = 7


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def make_weights(model):
    from oracle import synth
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = synth.fill_state_dict(shapes, WEIGHT_SEED, randomize_bn=True)
    return {k: torch.from_numpy(v) for k, v in sd.items()}, sd


def make_windows(n, seed):
    """Seeded synthetic 16 kHz audio cut into 1.5 s windows: band-limited noise bursts with a
    noise floor (cheap to generate for thousands of windows; the parity tests use the FM-speaker
    meeting)."""
    rng = np.random.default_rng([seed, 0xBE7C])
    x = rng.standard_normal((n, N_SAMPLES), dtype=np.float32)
    env = 0.05 + 0.15 * np.abs(np.sin(np.linspace(0, 9.0, N_SAMPLES, dtype=np.float32)))[None, :]
    return (x * env).astype(np.float32)


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
        # map the torch device index to NVML through the UUID (CUDA_VISIBLE_DEVICES may remap)
        try:
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


# ----------------------------------------------------------------------------- CPU oracle leg
def cpu_reference(seconds_budget, batch=64):
    """The oracle port of the path (numpy fbank + torch fp32 CAM++) on all host cores, on a
    bounded sample of the same workload."""
    from oracle import campplus_oracle, fbank_oracle
    import b200spk
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = b200spk.CAMPPlus(embedding_size=EMB)          # parameter shapes only (never run on CPU)
    _, sd = make_weights(model)
    wavs = make_windows(batch, seed=99)
    fbank_oracle.fbank_batch(wavs[:2])
    campplus_oracle.forward(sd, fbank_oracle.fbank_batch(wavs[:2]))
    done, t0 = 0, time.perf_counter()
    while True:
        feats = fbank_oracle.fbank_batch(wavs)
        campplus_oracle.forward(sd, feats)
        done += batch
        el = time.perf_counter() - t0
        if el >= seconds_budget:
            break
    return dict(value=done / el, unit="embeddings/s", cores=cores, kind="port",
                sample="%d x 1.5 s windows in batches of %d (%.1f s of CPU work)" % (done, batch, el))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
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
    tot_n, tot_t = 0, 0.0
    for _ in range(max(1, min(args.steps, 12))):
        n, t = step()
        tot_n, tot_t = tot_n + n, tot_t + t
    val = tot_n / tot_t
    line = {
        "impl": "reference", "metric": "CAM++ embeddings/sec (1.5 s windows)", "value": val, "unit": "embeddings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),        # the same workload key as the b200 arm; the CPU batching is in `sample`
        "cpu_baseline": {"value": val, "unit": "embeddings/s", "cores": cores, "kind": "port",
                         "sample": "%d x 1.5 s windows per ~%.0f s step, batches of %d" % (tot_n // max(1, args.steps), per_step, batch)},
        "e2e": {"value": val, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, batch):
    return {"workload": "CAM++ (7.2M, emb 512, random-init) embedding extraction: fbank+CMN -> forward on synthetic "
                        "16 kHz 1.5 s windows (BASELINE config 0 shape)",
            "segments_per_gpu_per_step": args.segments, "batch": batch, "precision": args.precision,
            "parallelism": "dp%d (sub-segment sharding, no collective)" % args.gpus,
            "l2": "inputs larger than L2 (%.0f MB of windows per step)" % (args.segments * N_SAMPLES * 4 / 1e6)}


# ----------------------------------------------------------------------------- diarization leg
def diarization_leg(args, dev, rank, world, dist):
    """BASELINE config 3: a synthetic 1-hour meeting -> 4799 sub-segments -> CAM++ (192-d) ->
    cosine affinity + spectral clustering, sub-segments sharded over the ranks, one all_gather of
    the embeddings.  Timed end to end from the waveform in pinned host memory to labels on the
    host; RTF = wall / audio seconds."""
    import b200spk
    from oracle import cluster_oracle, synth
    secs, K = args.meeting_seconds, args.meeting_speakers
    n = int(round(secs * 16000))
    if rank == 0:
        wav_np, turns = synth.fm_meeting(secs, K, seed=18)
        wav = torch.from_numpy(wav_np).pin_memory()
    else:
        wav, turns = torch.empty(n, dtype=torch.float32).pin_memory(), None
    if dist is not None:
        tmp = wav.to(dev)
        dist.broadcast(tmp, src=0)
        wav.copy_(tmp)
        del tmp
    torch.manual_seed(1)                                   # default init, BN not randomised
    model = b200spk.CAMPPlus(embedding_size=192, precision=args.precision).to(dev).eval()
    dz = b200spk.Diarizer(b200spk.FBank(80, 16000, mean_nor=True), model,
                          b200spk.SpectralCluster(min_num_spks=1, max_num_spks=15, pval=0.012, device=dev),
                          device=dev, batchsize=args.batch)
    times = []
    for it in range(4):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        np.random.seed(0)
        t0 = time.perf_counter()
        chunks, labels = dz(wav)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([sorted(times[1:])[1]], device=dev)    # median of the 3 timed runs
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    if rank != 0:
        return None
    truth, pure = synth.turn_labels(chunks, turns)
    purity = sum(np.bincount(truth[pure & (labels == c)]).max() for c in np.unique(labels[pure])) / pure.sum()
    # parity of the back end at full size: the CPU oracle on the SAME embeddings
    with torch.no_grad():
        emb = b200spk.gather_embeddings(dz.extract(wav.to(dev), chunks), len(chunks)).cpu().numpy() if world == 1 else None
    parity = None
    if emb is not None:
        np.random.seed(0)
        ref, st = cluster_oracle.spectral_cluster(emb, 1, 15, 0.012, return_stages=True)
        m = cluster_oracle.match_labels(ref, labels)
        parity = {"k_oracle": int(st["k"]), "mismatch_pure": int((m != ref)[pure].sum()), "mismatch_total": int((m != ref).sum())}
    return {"audio_seconds": secs, "n_subsegments": len(chunks), "speakers": K, "wall_s": wall, "rtf": wall / secs,
            "k": int(dz.cluster.last["k"]), "purity_on_single_speaker_segments": float(purity),
            "krylov_dim": int(dz.cluster.last["krylov"]), "vs_oracle_backend": parity}


# ----------------------------------------------------------------------------- GPU leg
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--segments", type=int, default=16384, help="1.5 s windows per GPU per step")
    ap.add_argument("--batch", type=int, default=2048, help="windows per fbank/forward call")
    ap.add_argument("--chunk", type=int, default=0, help="coarse sub-batch (D-TDNN part), 0 = auto")
    ap.add_argument("--fine", type=int, default=0, help="fine sub-batch (2-D front), 0 = auto")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-meeting", action="store_true", help="skip the 1-hour diarization leg")
    ap.add_argument("--meeting-seconds", type=float, default=3600.0)
    ap.add_argument("--meeting-speakers", type=int, default=8)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import b200spk
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    chunk = None
    if args.chunk or args.fine:
        chunk = (args.chunk or 2048, args.fine or min(args.chunk or 2048, 1024))
    model = b200spk.CAMPPlus(embedding_size=EMB, precision=args.precision, chunk=chunk)
    tsd, _ = make_weights(model)
    model.load_state_dict(tsd)
    model = model.to(dev).eval()
    fb = b200spk.FBank(80, 16000, mean_nor=True)
    ex = b200spk.EmbeddingExtractor(fb, model, device=dev, batchsize=args.batch, reuse_output=True)

    S = args.segments
    host = torch.from_numpy(make_windows(S, seed=1000 + rank)).pin_memory()
    wav_dev = host.to(dev)

    L = b200spk.lib()
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
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

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device(None)
    barrier()
    launches0 = L.spk_launch_count()
    recs = []
    with ClockSampler(local) as clk:
        t_start, t_end = ev(), ev()
        t_start.record()
        for _ in range(args.steps):
            step_device(recs)
        t_end.record()
        barrier()
    launches = L.spk_launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    for e0, e1, e2 in recs:
        fb_ms.append(e0.elapsed_time(e1))
        fw_ms.append(e1.elapsed_time(e2))
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * S * args.steps / (ms / 1e3)

    # ---- end to end through the public API with pinned host buffers
    for _ in range(2):
        ex(host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        emb_host = ex(host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = world * S * args.steps / e2e_s
    assert emb_host.shape == (S, EMB) and bool(torch.isfinite(emb_host).all())

    meeting = None if args.no_meeting else diarization_leg(args, dev, rank, world, dist)

    if rank == 0:
        pk = peaks()
        calls = len(fw_ms)
        fw_avg = sum(fw_ms) / calls            # ms per forward call (args.batch segments)
        fb_avg = sum(fb_ms) / calls
        segs_per_call = S / (S // args.batch + (1 if S % args.batch else 0))
        tf = GFLOP_PER_SEG * segs_per_call / fw_avg          # GFLOP/ms == TFLOP/s
        gbs = FBANK_BYTES_PER_SEG * segs_per_call / fb_avg / 1e6
        # DRAM bytes of one forward call (ncu, profiles/r01_traffic_o.md: 31.29 GB per 2048 segments in bf16 mode)
        traffic = DRAM_BYTES_PER_SEG_BF16 * segs_per_call if args.precision == "bf16" else None
        line = {
            "metric": "CAM++ embeddings/sec (1.5 s windows)", "value": value, "unit": "embeddings/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, args.batch),
            "e2e": {"value": e2e, "unit": "embeddings/s", "h2d_bytes_per_step": S * N_SAMPLES * 4,
                    "d2h_bytes_per_step": S * EMB * 4},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "CAM++ conv stack (all implicit-GEMM launches of one forward call)",
                         "bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": tf / pk["bf16_sustained"], "traffic": traffic, "peak_source": pk["source"] + " (sustained)",
                         "flops_per_launch": GFLOP_PER_SEG * 1e9 * segs_per_call, "ms_per_launch": fw_avg,
                         "traffic_source": "ncu dram__bytes_read+write over one forward call, profiles/r01_traffic_o.md"},
            # the same forward call against HBM: 15.3 MB of DRAM traffic per segment (measured) is the tighter bound
            "roofline_hbm": {"kernel": "CAM++ forward call (all launches)", "bound": "hbm",
                             "achieved": (traffic / (fw_avg * 1e-3) / 1e9) if traffic else None, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": (traffic / (fw_avg * 1e-3) / 1e9 / pk["hbm"]) if traffic else None, "traffic": traffic},
            "roofline_fbank": {"kernel": "fbank_kernel<fused>", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"],
                               "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None,
                               "bytes_per_launch": FBANK_BYTES_PER_SEG * segs_per_call, "ms_per_launch": fb_avg},
            "clocks": clk.summary(),
        }
        if meeting is not None:
            line["diarization"] = meeting
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_reference(args.cpu_seconds)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
