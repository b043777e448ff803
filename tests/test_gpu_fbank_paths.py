"""GPU: the fused-CMN fbank path and the frame-range + second-pass path must agree."""
import numpy as np
import pytest
import torch

import b200spk
from oracle import fbank_oracle, synth

pytestmark = pytest.mark.gpu


def test_fused_and_two_pass_paths_agree_3s():
    # m = 298: batch >= 2*SMs takes the fused path, a single utterance the two-pass path
    wavs = synth.white_noise(400, 48000, seed=90)
    x = torch.from_numpy(wavs).cuda()
    fused_raw = b200spk.fbank_batch(x, 80, False)
    single_raw = torch.cat([b200spk.fbank_batch(x[i:i + 1], 80, False) for i in (0, 7, 399)])
    assert torch.equal(fused_raw[[0, 7, 399]], single_raw)
    fused = b200spk.fbank_batch(x, 80, True)
    single = torch.cat([b200spk.fbank_batch(x[i:i + 1], 80, True) for i in (0, 7, 399)])
    assert torch.equal(fused[[0, 7, 399]], single)


def test_every_element_close_to_fp64_truth_strict():
    # strict elementwise check, no outlier allowance
    wavs = synth.white_noise(64, 24000, seed=91)
    got = b200spk.fbank_batch(torch.from_numpy(wavs).cuda(), 80, True).cpu().numpy()
    ref64 = fbank_oracle.fbank_batch(wavs, dtype=np.float64)
    err = np.abs(got - ref64)
    assert err.max() <= 1e-4, err.max()         # north-star tolerance on every element
    assert err.mean() <= 2e-6
