"""Kaldi ``ark,scp`` output for bulk extraction: the ``WriteHelper('ark,scp:...')`` of
``speakerlab/bin/extract.py:73-96`` (kaldiio is not a dependency here; the binary archive format is small).

Binary float matrix entry:  ``<key> \\0B FM \\4 <rows:int32> \\4 <cols:int32> <rows*cols float32 LE>``; float vector:
``<key> \\0B FV \\4 <dim:int32> <data>``.  The scp line is ``<key> <ark path>:<offset of \\0B>``.  The reference
writes every embedding as a [1, E] matrix (the model output of one utterance), so that is the default here too.
"""
import os
import struct

import numpy as np


class ArkWriter:
    """``with ArkWriter(ark_path, scp_path) as w: w(key, array)`` - same call convention as kaldiio's WriteHelper."""

    def __init__(self, ark_path, scp_path=None):
        self.ark_path, self.scp_path = ark_path, scp_path
        self._ark = open(ark_path, "wb")
        self._scp = open(scp_path, "w") if scp_path else None

    def __call__(self, key, array):
        a = np.ascontiguousarray(np.asarray(array, dtype="<f4"))
        assert a.ndim in (1, 2) and " " not in key
        self._ark.write(key.encode() + b" ")
        offset = self._ark.tell()
        if a.ndim == 2:
            self._ark.write(b"\0BFM \4" + struct.pack("<i", a.shape[0]) + b"\4" + struct.pack("<i", a.shape[1]))
        else:
            self._ark.write(b"\0BFV \4" + struct.pack("<i", a.shape[0]))
        self._ark.write(a.tobytes())
        if self._scp:
            self._scp.write("%s %s:%d\n" % (key, os.path.abspath(self.ark_path), offset))

    def write_batch(self, keys, embeddings):
        """keys: list of str; embeddings: [N, E] host array or CPU tensor -> one [1, E] matrix per key."""
        emb = np.asarray(embeddings, dtype=np.float32)
        for k, e in zip(keys, emb):
            self(k, e[None, :])

    def close(self):
        self._ark.close()
        if self._scp:
            self._scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_ark(path):
    """Generator of (key, array) over a binary float ark (for tests and for loading extracted embeddings back)."""
    with open(path, "rb") as f:
        while True:
            key = bytearray()
            while True:
                ch = f.read(1)
                if not ch:
                    return
                if ch == b" ":
                    break
                key += ch
            yield key.decode(), _read_entry(f)


def _read_entry(f):
    assert f.read(2) == b"\0B", "not a binary Kaldi entry"
    kind = f.read(3)
    if kind == b"FM ":
        assert f.read(1) == b"\4"
        rows = struct.unpack("<i", f.read(4))[0]
        assert f.read(1) == b"\4"
        cols = struct.unpack("<i", f.read(4))[0]
        return np.frombuffer(f.read(4 * rows * cols), dtype="<f4").reshape(rows, cols).copy()
    assert kind == b"FV ", kind
    assert f.read(1) == b"\4"
    dim = struct.unpack("<i", f.read(4))[0]
    return np.frombuffer(f.read(4 * dim), dtype="<f4").copy()


def read_scp(path):
    """{key: array} from an scp file (each line ``key ark_path:offset``)."""
    out = {}
    for line in open(path):
        key, loc = line.split()
        ark, off = loc.rsplit(":", 1)
        with open(ark, "rb") as f:
            f.seek(int(off))
            out[key] = _read_entry(f)
    return out
