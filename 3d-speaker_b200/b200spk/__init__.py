"""b200spk - B200-native (sm_100a) implementation of 3D-Speaker's embedding-extraction hot path.

Python mirrors of the reference call sites over the C ABI of ``libb200spk.so``:

    FBank            <- speakerlab.process.processor.FBank
    CAMPPlus         <- speakerlab.models.campplus.DTDNN.CAMPPlus
    ERes2NetV2       <- speakerlab.models.eres2net.ERes2NetV2.ERes2NetV2
    ERes2Net, ERes2Net_huge <- speakerlab.models.eres2net.ERes2Net.ERes2Net / ERes2Net_huge.ERes2Net
    ECAPA_TDNN       <- speakerlab.models.ecapa_tdnn.ECAPA_TDNN.ECAPA_TDNN
    SpectralCluster  <- speakerlab.process.cluster.SpectralCluster
    AHCluster, CommonClustering <- speakerlab.process.cluster.{AHCluster, CommonClustering}
    EmbeddingExtractor: the batched fbank -> model loop of
                     speakerlab/bin/infer_diarization.py:621-639 with host buffers in/out

There is no CPU fallback: without the built library import of the compute classes fails, and
without an sm_100 device every compute call raises ``SpkError``.
"""
from ._lib import SpkError, lib, LIB_PATH  # noqa: F401
from .fbank import FBank, fbank_batch, fbank_windows, num_frames  # noqa: F401
from .campplus import CAMPPlus  # noqa: F401
from .eres2netv2 import ERes2NetV2  # noqa: F401
from .eres2net import ERes2Net, ERes2Net_huge  # noqa: F401
from .ecapa_tdnn import ECAPA_TDNN  # noqa: F401
from .cluster import SpectralCluster, AHCluster, CommonClustering, cosine_pairs  # noqa: F401
from .extract import EmbeddingExtractor  # noqa: F401
from .bulk import BulkExtractor, chunk_table, segment_mean  # noqa: F401
from .postprocess import compress_segments, rttm_lines, write_rttm  # noqa: F401
from .kaldi_io import ArkWriter, read_ark, read_scp  # noqa: F401
from .metrics import det_curve, eer_min_dcf  # noqa: F401
from .diarize import Diarizer, cut_windows, gather_embeddings, shard_range  # noqa: F401

__all__ = ["FBank", "CAMPPlus", "ERes2NetV2", "ERes2Net", "ERes2Net_huge", "ECAPA_TDNN", "SpectralCluster", "AHCluster", "CommonClustering", "cosine_pairs", "EmbeddingExtractor", "Diarizer", "SpkError", "fbank_batch", "num_frames", "lib"]
