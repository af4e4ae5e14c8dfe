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


# ---------------------------------------------------------------------------------------------- Kaldi <Plda> models
def _read_plda_vec_binary(f):
    tok = f.read(3)
    if tok not in (b"FV ", b"DV "):
        raise ValueError(f"Kaldi <Plda>: bad vector type {tok!r}")
    if f.read(1) != b"\4":
        raise ValueError("Kaldi <Plda>: bad vector header")
    dim = struct.unpack("<i", f.read(4))[0]
    if tok == b"FV ":
        return np.frombuffer(f.read(4 * dim), dtype="<f4")
    return np.frombuffer(f.read(8 * dim), dtype="<f8")


def _read_plda_mat_binary(f):
    tok = f.read(3)
    if tok not in (b"FM ", b"DM "):
        raise ValueError(f"Kaldi <Plda>: unsupported matrix type {tok!r} (FM / DM expected; compressed and sparse "
                         "matrices never occur in ivector-compute-plda output)")
    hdr = f.read(10)
    if hdr[0:1] != b"\4" or hdr[5:6] != b"\4":
        raise ValueError("Kaldi <Plda>: bad matrix header")
    rows, cols = struct.unpack("<i", hdr[1:5])[0], struct.unpack("<i", hdr[6:10])[0]
    if tok == b"FM ":
        return np.frombuffer(f.read(4 * rows * cols), dtype="<f4").reshape(rows, cols)
    return np.frombuffer(f.read(8 * rows * cols), dtype="<f8").reshape(rows, cols)


def _read_ascii_mat(f):
    """Rows of a Kaldi text matrix after the opening ' [' was consumed: one row per line, the last ends with ']'."""
    rows = []
    while True:
        line = f.readline()
        if not line:
            raise ValueError("Kaldi <Plda>: unterminated text matrix")
        line = line.strip()
        if not line:
            continue
        last = line.endswith(b"]")
        vals = line.rstrip(b"]").split()
        if vals:
            rows.append(np.array(vals, dtype=np.float32))   # kaldi_io._read_mat_ascii parses into float32
        if last:
            return np.vstack(rows)


def read_plda(path_or_fd):
    """Kaldi `<Plda>` model (binary or text) -> (mean, transform, psi), same contract as
    `wespeaker/utils/plda/kaldi_utils.py:24-55` (whose vector / matrix readers are `:58-108`)."""
    f = open(path_or_fd, "rb") if isinstance(path_or_fd, (str, os.PathLike)) else path_or_fd
    try:
        binary = f.read(2)
        if binary == b"\0B":
            if f.read(7) != b"<Plda> ":
                raise ValueError("Kaldi <Plda>: missing <Plda> token")
            mean = _read_plda_vec_binary(f)
            trans = _read_plda_mat_binary(f)
            psi = _read_plda_vec_binary(f)
        else:
            if binary + f.read(5) != b"<Plda> ":
                raise ValueError("Kaldi <Plda>: missing <Plda> token")
            mean = np.array(f.readline().decode().strip(" \n[]").split(), dtype=float)
            if f.read(2) != b" [":
                raise ValueError("Kaldi <Plda>: expected a text matrix after the mean")
            trans = _read_ascii_mat(f)
            psi = np.array(f.readline().decode().strip(" \n[]").split(), dtype=float)
        if f.read(8) != b"</Plda> ":
            raise ValueError("Kaldi <Plda>: missing </Plda> token")
    finally:
        if f is not path_or_fd:
            f.close()
    return mean, trans, psi


def write_plda(path, mean, transform, psi, binary=True):
    """Write a Kaldi `<Plda>` model (double precision, like ivector-compute-plda): the inverse of read_plda."""
    mean, transform, psi = (np.asarray(a, dtype=np.float64) for a in (mean, transform, psi))
    with open(path, "wb") as f:
        if binary:
            f.write(b"\0B<Plda> ")
            f.write(b"DV \4" + struct.pack("<i", mean.shape[0]) + mean.astype("<f8").tobytes())
            f.write(b"DM \4" + struct.pack("<i", transform.shape[0]) + b"\4" + struct.pack("<i", transform.shape[1])
                    + np.ascontiguousarray(transform, dtype="<f8").tobytes())
            f.write(b"DV \4" + struct.pack("<i", psi.shape[0]) + psi.astype("<f8").tobytes())
            f.write(b"</Plda> ")
        else:
            fmt = lambda v: " ".join(repr(float(x)) for x in v)  # noqa: E731
            f.write(("<Plda>  [ " + fmt(mean) + " ]\n").encode())
            f.write(b" [\n")
            for i, row in enumerate(transform):
                f.write(("  " + fmt(row) + (" ]\n" if i == transform.shape[0] - 1 else "\n")).encode())
            f.write((" [ " + fmt(psi) + " ]\n").encode())
            f.write(b"</Plda> ")


# ---------------------------------------------------------------------------------------------- Kaldi matrices
def _read_mat_binary(f):
    """One binary Kaldi matrix after the '\\0B' marker: FM / DM (plain float / double) or CM / CM2 / CM3 (Kaldi's
    compressed-matrix formats, compressed-matrix.h): what `kaldiio.load_mat` returns for `dataset/processor.py:190`."""
    tok = f.read(3)
    if tok in (b"FM ", b"DM "):
        hdr = f.read(10)
        if hdr[0:1] != b"\4" or hdr[5:6] != b"\4":
            raise ValueError("bad Kaldi matrix header")
        rows, cols = struct.unpack("<i", hdr[1:5])[0], struct.unpack("<i", hdr[6:10])[0]
        if tok == b"FM ":
            return np.frombuffer(f.read(4 * rows * cols), dtype="<f4").reshape(rows, cols).copy()
        return np.frombuffer(f.read(8 * rows * cols), dtype="<f8").reshape(rows, cols).astype(np.float32)
    if tok in (b"CM ", b"CM2", b"CM3"):
        if tok != b"CM ":
            f.read(1)                                     # the space after the 3-character token
        vmin, vrange, rows, cols = struct.unpack("<ffii", f.read(16))
        if tok == b"CM2":
            d = np.frombuffer(f.read(2 * rows * cols), dtype="<u2").reshape(rows, cols)
            return (vmin + vrange * (d.astype(np.float32) / 65535.0)).astype(np.float32)
        if tok == b"CM3":
            d = np.frombuffer(f.read(rows * cols), dtype="u1").reshape(rows, cols)
            return (vmin + vrange * (d.astype(np.float32) / 255.0)).astype(np.float32)
        # CM: per-column percentile headers (4 x uint16) then column-major bytes, piecewise-linear in three segments
        ph = np.frombuffer(f.read(8 * cols), dtype="<u2").reshape(cols, 4).astype(np.float32)
        p = vmin + vrange * (ph / 65535.0)                # (cols, 4): 0th, 25th, 75th, 100th percentile values
        d = np.frombuffer(f.read(rows * cols), dtype="u1").reshape(cols, rows).astype(np.float32)
        p0, p25, p75, p100 = (p[:, i:i + 1] for i in range(4))
        out = np.where(d <= 64, p0 + (p25 - p0) * d / 64.0,
                       np.where(d <= 192, p25 + (p75 - p25) * (d - 64.0) / 128.0, p75 + (p100 - p75) * (d - 192.0) / 63.0))
        return np.ascontiguousarray(out.T.astype(np.float32))
    raise ValueError(f"unsupported Kaldi matrix type {tok!r}")


def load_mat(location: str) -> np.ndarray:
    """`kaldiio.load_mat("file.ark:offset")` (or a file holding one matrix): (T, F) float32."""
    path, sep, off = location.rpartition(":")
    if not sep or not off.isdigit():
        path, off = location, None
    with open(path, "rb") as f:
        if off is not None:
            f.seek(int(off))
        else:                                             # a key may precede the binary marker in a bare ark
            head = f.read(2)
            if head != b"\0B":
                f.seek(0)
                while f.read(1) not in (b" ", b""):
                    pass
                head = f.read(2)
            f.seek(f.tell() - 2)
        if f.read(2) != b"\0B":
            raise ValueError(f"{location}: not a binary Kaldi object")
        return _read_mat_binary(f)


class MatrixWriter:
    """``with MatrixWriter(ark, scp) as w: w(key, mat)`` — float matrices ('FM '), the feature ark/scp of Kaldi recipes."""

    def __init__(self, ark_path: str, scp_path: str | None = None):
        self.ark_path = os.path.abspath(ark_path)
        self.ark = open(self.ark_path, "wb")
        self.scp = open(scp_path, "w") if scp_path else None

    def __call__(self, key: str, mat):
        m = np.ascontiguousarray(np.asarray(mat), dtype=np.float32)
        self.ark.write(key.encode() + b" ")
        off = self.ark.tell()
        self.ark.write(b"\0BFM \4" + struct.pack("<i", m.shape[0]) + b"\4" + struct.pack("<i", m.shape[1]) + m.tobytes())
        if self.scp:
            self.scp.write(f"{key} {self.ark_path}:{off}\n")
        return f"{self.ark_path}:{off}"

    def close(self):
        self.ark.close()
        if self.scp:
            self.scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
