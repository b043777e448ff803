"""Batched sub-segment embedding extraction: the hot loop of
``Diarization3Dspeaker.do_emb_extraction`` (speakerlab/bin/infer_diarization.py:621-639) and
``local/extract_diar_embeddings.py:187-203``: per batch H2D of the waveforms, batched fbank,
network forward, D2H of the embeddings - with host buffers on both sides.
"""
import torch


class EmbeddingExtractor:
    def __init__(self, feature_extractor, embedding_model, device="cuda:0", batchsize=64, reuse_output=False, head=512):
        """``reuse_output``: return a cached pinned host buffer (allocating pinned memory costs milliseconds per call);
        the result is then only valid until the next call.  ``head``: size of the first sub-batch - nothing overlaps
        the first host-to-device copy, so it is kept short; the following batches grow by 4x per step (the copy of
        batch i+1 hides behind the compute of batch i only if it is not much larger) until they reach ``batchsize``."""
        self.feature_extractor = feature_extractor
        self.embedding_model = embedding_model
        self.device = torch.device(device)
        self.batchsize = batchsize
        self.reuse_output = reuse_output
        self.head = head
        self._copy_stream = None
        self._out = None

    def __call__(self, wavs):
        """wavs: host tensor [N, L] (pinned for async copies) or [N, 1, L] -> host [N, E].

        Same per-batch H2D / fbank / forward / D2H sequence as the reference loop; the copy of batch
        i+1 runs on a second stream while batch i computes (the reference serialises them)."""
        if wavs.dim() == 3:
            wavs = wavs[:, 0, :]
        N = wavs.shape[0]
        if N == 0:
            return torch.empty((0, self.embedding_model.embedding_size))
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        size = min(self.batchsize, self.head) if (self.head and N > self.batchsize) else self.batchsize
        bounds = [0, min(size, N)]
        while bounds[-1] < N:
            size = min(self.batchsize, size * 4)
            bounds.append(min(bounds[-1] + size, N))
        starts = bounds[:-1]
        ends = dict(zip(bounds[:-1], bounds[1:]))

        def stage(st):
            with torch.cuda.stream(cs):
                buf = wavs[st:ends[st]].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            return buf, ev

        out = None
        nxt = stage(starts[0])
        with torch.no_grad():
            for i, st in enumerate(starts):
                wb, ev = nxt
                if i + 1 < len(starts):
                    nxt = stage(starts[i + 1])
                main.wait_event(ev)
                wb.record_stream(main)
                feats = self.feature_extractor.batch(wb)
                emb = self.embedding_model(feats)
                if out is None:
                    if self.reuse_output and self._out is not None and self._out.shape == (N, emb.shape[1]):
                        out = self._out
                    else:
                        out = torch.empty((N, emb.shape[1]), dtype=torch.float32, pin_memory=wavs.is_pinned())
                        if self.reuse_output:
                            self._out = out
                out[st:st + emb.shape[0]].copy_(emb, non_blocking=True)
        main.synchronize()
        return out

    def extract_device(self, wavs_dev):
        """Same loop with inputs already resident in HBM; returns device embeddings."""
        outs = []
        with torch.no_grad():
            for st in range(0, wavs_dev.shape[0], self.batchsize):
                feats = self.feature_extractor.batch(wavs_dev[st:st + self.batchsize])
                outs.append(self.embedding_model(feats))
        return torch.cat(outs, dim=0)
