"""CPU: the C-ABI library loads, exports every symbol include/b200spk.h declares, and fails
loudly (no CPU fallback) when there is no device.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

import b200spk
from b200spk import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200spk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.exported_names()) == names


def test_abi_version_and_frame_count():
    L = _lib.lib()
    assert L.spk_abi_version() == 1
    assert L.spk_fbank_num_frames(24000) == 148
    assert L.spk_fbank_num_frames(48000) == 298
    assert L.spk_fbank_num_frames(399) == 0
    assert L.spk_fbank_num_frames(400) == 1


def test_struct_layout_matches_header():
    # spk_op_t is 41 int32 + 2 float + phase/reserved; spk_buf_t is int64 + 2 int32
    assert C.sizeof(_lib.SpkBuf) == 16
    assert C.sizeof(_lib.SpkOp) == 4 * (1 + 3 + 3 + 3 + 2 + 6 + 8 + 1 + 3 + 3 + 4 + 4) + 8 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_cpu_fallback():
    L = _lib.lib()
    assert L.spk_device_check(0) == -5
    assert b"no CPU fallback" in L.spk_last_error()
    wav = torch.zeros(1, 24000)
    with pytest.raises(b200spk.SpkError) as e:
        b200spk.fbank_batch(wav)          # host entry point still needs the device
    assert e.value.code == -5
    model = b200spk.CAMPPlus(embedding_size=192)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 148, 80))


def test_invalid_arguments_are_rejected_before_any_device_work():
    L = _lib.lib()
    buf = torch.zeros(8)
    # too short for one frame: the reference trips kaldi.py:142
    rc = L.spk_fbank_f32(C.c_void_p(buf.data_ptr()), 1, 100, 100, C.c_void_p(buf.data_ptr()), 80, 1, None)
    assert rc == -1 and b"window size" in L.spk_last_error()
    rc = L.spk_fbank_f32(None, 1, 24000, 24000, None, 80, 1, None)
    assert rc == -1


def test_state_dict_layout_is_the_reference_layout(golden_dir):
    import json
    lay = json.load(open(os.path.join(golden_dir, "state_dict_layouts.json")))
    for emb in (192, 512):
        sd = b200spk.CAMPPlus(embedding_size=emb).state_dict()
        ref = lay["campplus_e%d" % emb]
        assert list(sd.keys()) == list(ref.keys())
        assert all(list(sd[k].shape) == ref[k] for k in ref)


def test_ecapa_state_dict_layout_is_the_reference_layout(golden_dir):
    import json
    lay = json.load(open(os.path.join(golden_dir, "state_dict_layouts.json")))
    for c in (512, 1024):
        sd = b200spk.ECAPA_TDNN(80, channels=[c, c, c, c, 3 * c]).state_dict()
        ref = lay["ecapa_c%d" % c]
        assert len(ref) == 231
        assert list(sd.keys()) == list(ref.keys())
        assert all(list(sd[k].shape) == ref[k] for k in ref)


def test_eres2net_v1_state_dict_layout_is_the_reference_layout(golden_dir):
    import json
    lay = json.load(open(os.path.join(golden_dir, "state_dict_layouts.json")))
    for variant, model in (("base", b200spk.ERes2Net()), ("large", b200spk.ERes2Net(m_channels=64)), ("huge", b200spk.ERes2Net_huge())):
        sd = model.state_dict()
        ref = lay["eres2net_" + variant]
        assert list(sd.keys()) == list(ref.keys())
        assert all(list(sd[k].shape) == ref[k] for k in ref)


def test_fbank_mirror_interface():
    fb = b200spk.FBank(80, 16000, mean_nor=True)
    assert (fb.n_mels, fb.sample_rate, fb.mean_nor) == (80, 16000, True)
    with pytest.raises(AssertionError):
        b200spk.FBank(80, 8000)(torch.zeros(16000))
    with pytest.raises(AssertionError):
        fb(torch.zeros(1, 100))
