"""Batched sub-segment embedding extraction: the hot loop of
``Diarization3Dspeaker.do_emb_extraction`` (speakerlab/bin/infer_diarization.py:621-639) and
``local/extract_diar_embeddings.py:187-203``: per batch H2D of the waveforms, batched fbank,
network forward, D2H of the embeddings - with host buffers on both sides.
"""
import torch


class EmbeddingExtractor:
    def __init__(self, feature_extractor, embedding_model, device="cuda:0", batchsize=64):
        self.feature_extractor = feature_extractor
        self.embedding_model = embedding_model
        self.device = torch.device(device)
        self.batchsize = batchsize
        self._copy_stream = None

    def __call__(self, wavs):
        """wavs: host tensor [N, L] (pinned for async copies) or [N, 1, L] -> host [N, E]."""
        if wavs.dim() == 3:
            wavs = wavs[:, 0, :]
        N = wavs.shape[0]
        out = None
        with torch.no_grad():
            for st in range(0, N, self.batchsize):
                wb = wavs[st:st + self.batchsize].to(self.device, non_blocking=True)
                feats = self.feature_extractor.batch(wb)
                emb = self.embedding_model(feats)
                if out is None:
                    out = torch.empty((N, emb.shape[1]), dtype=torch.float32,
                                      pin_memory=wavs.is_pinned())
                out[st:st + emb.shape[0]].copy_(emb, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out if out is not None else torch.empty((0, 0))

    def extract_device(self, wavs_dev):
        """Same loop with inputs already resident in HBM; returns device embeddings."""
        outs = []
        with torch.no_grad():
            for st in range(0, wavs_dev.shape[0], self.batchsize):
                feats = self.feature_extractor.batch(wavs_dev[st:st + self.batchsize])
                outs.append(self.embedding_model(feats))
        return torch.cat(outs, dim=0)
