#!/usr/bin/env python
"""Goldens from the reference's OWN C++ code (oracle/_ref/libref_engine.so, built by `make -C oracle` from
/root/reference/runtime/core/{speaker/speaker_engine.cc, frontend/feature_pipeline.cc, frontend/fft.cc, frontend/fbank.h}):

* full-mode features = the native fbank twin (`fbank.h:138-198`) of the torchaudio fbank the Python path uses,
* chunk-mode features = `SpeakerEngine::ExtractFeature` (`speaker_engine.cc:77-139`) incl. head-padding of the last chunk,
  and the same after `ApplyMean` (`:62-75`),
* `SpeakerEngine::CosineSimilarity` (`:161-172`).

Stored as INDICES where possible (which full-mode frame each chunk row is a bit-copy of), so the fixture stays small."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_ref():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_engine.so"))
    L.ref_extract_feature.restype = C.c_int
    L.ref_extract_feature.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int)]
    L.ref_cosine_similarity.restype = C.c_float
    L.ref_cosine_similarity.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ref_free.argtypes = [C.POINTER(C.c_float)]
    return L


def ref_features(L, pcm_i16, samples_per_chunk, apply_mean):
    out, shape = C.POINTER(C.c_float)(), (C.c_int * 3)()
    pcm = np.ascontiguousarray(pcm_i16, dtype=np.int16)
    nc = L.ref_extract_feature(pcm.ctypes.data, len(pcm), samples_per_chunk, apply_mean, C.byref(out), shape)
    assert nc >= 0
    a = np.ctypeslib.as_array(out, shape=(max(1, shape[0]) * max(1, shape[1]) * max(1, shape[2]),)).copy()
    L.ref_free(out)
    return a[: shape[0] * shape[1] * shape[2]].reshape(shape[0], shape[1], shape[2])


def main():
    from wespeaker_b200 import synthetic as syn
    L = load_ref()
    out = {}
    for nsamp in (9000, 20000, 32000, 50000, 80333, 163840):
        pcm = syn.make_wavs(1, nsamp, seed=7)[0].astype(np.int16)
        full = ref_features(L, pcm, 0, 0)[0]                       # (T, 80): every frame of the native fbank
        chunks = ref_features(L, pcm, 32000, 0)                    # (nc, 198, 80)
        chunks_cmn = ref_features(L, pcm, 32000, 1)
        # each chunk row must be a bit-copy of one full-mode frame: store the frame index
        idx = np.empty(chunks.shape[:2], dtype=np.int32)
        for c in range(chunks.shape[0]):
            for r in range(chunks.shape[1]):
                hit = np.nonzero((full == chunks[c, r]).all(axis=1))[0]
                assert len(hit) >= 1, (nsamp, c, r)
                # consecutive rows continue the previous index when several frames are identical (never with noise input)
                idx[c, r] = hit[0]
        out[f"chunk_index_{nsamp}"] = idx
        out[f"nframes_{nsamp}"] = np.array(full.shape[0])
        if nsamp in (9000, 50000):                                             # keep the fixture small
            out[f"full_{nsamp}"] = full.astype(np.float32)
            out[f"lastchunk_cmn_{nsamp}"] = chunks_cmn[-1].astype(np.float32)  # ApplyMean of the (padded) last chunk
        print(nsamp, "frames", full.shape[0], "chunks", chunks.shape[0])
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((6, 192)).astype(np.float32), rng.standard_normal((6, 192)).astype(np.float32)
    out["cos_a"], out["cos_b"] = a, b
    out["cos"] = np.array([L.ref_cosine_similarity(a[i].ctypes.data, b[i].ctypes.data, 192) for i in range(6)], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_engine.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
