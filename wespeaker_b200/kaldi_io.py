"""Minimal Kaldi ark/scp IO for float vectors — the on-disk format on both sides of the hot path
(`wespeaker/bin/extract.py:105-111,137-139` writes embeddings with kaldiio.WriteHelper('ark,scp:...');
`wespeaker/utils/plda/plda_utils.py:20-29` reads them).  kaldiio is not installed in this image, so the
byte format is restated here: ``key + ' ' + '\\0B' + 'FV ' + '\\4' + int32(dim) + float32[dim]``,
scp line = ``key path:offset`` with offset pointing at the '\\0B' marker."""
from __future__ import annotations

import os
import struct

import numpy as np


class VectorWriter:
    """``with VectorWriter(ark, scp) as w: w(key, vec)`` — same call shape as kaldiio.WriteHelper."""

    def __init__(self, ark_path: str, scp_path: str | None = None):
        self.ark_path = os.path.abspath(ark_path)
        self.ark = open(self.ark_path, "wb")
        self.scp = open(scp_path, "w") if scp_path else None

    def __call__(self, key: str, vec):
        v = np.ascontiguousarray(np.asarray(vec), dtype=np.float32).reshape(-1)
        self.ark.write(key.encode() + b" ")
        off = self.ark.tell()
        self.ark.write(b"\0BFV \4" + struct.pack("<i", v.shape[0]) + v.tobytes())
        if self.scp:
            self.scp.write(f"{key} {self.ark_path}:{off}\n")

    def close(self):
        self.ark.close()
        if self.scp:
            self.scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _read_vec(f):
    hdr = f.read(2)
    if hdr != b"\0B":
        raise ValueError("not a binary Kaldi object")
    tok = f.read(3)
    if tok not in (b"FV ", b"DV "):
        raise ValueError(f"unsupported Kaldi type {tok!r} (float/double vector expected)")
    if f.read(1) != b"\4":
        raise ValueError("bad vector header")
    dim = struct.unpack("<i", f.read(4))[0]
    if tok == b"FV ":
        return np.frombuffer(f.read(4 * dim), dtype="<f4").copy()
    return np.frombuffer(f.read(8 * dim), dtype="<f8").astype(np.float32)


def load_scp_sequential(scp_path: str):
    """Yield (key, vector) like kaldiio.load_scp_sequential."""
    handles = {}
    try:
        with open(scp_path) as fin:
            for line in fin:
                line = line.strip()
                if not line:
                    continue
                key, loc = line.split(None, 1)
                path, _, off = loc.rpartition(":")
                if path not in handles:
                    handles[path] = open(path, "rb")
                f = handles[path]
                f.seek(int(off))
                yield key, _read_vec(f)
    finally:
        for f in handles.values():
            f.close()


def read_vec_scp_file(scp_file: str):
    """`wespeaker/utils/plda/plda_utils.py:20-29`."""
    return {k: v for k, v in load_scp_sequential(scp_file)}


def load_ark(ark_path: str):
    with open(ark_path, "rb") as f:
        while True:
            key = b""
            while True:
                c = f.read(1)
                if not c:
                    return
                if c == b" ":
                    break
                key += c
            yield key.decode(), _read_vec(f)
