"""Chunk-and-average long-audio mode: wespeaker_b200.speaker_engine against the literal restatement
of runtime/core/speaker/speaker_engine.cc in oracle/speaker_engine_py.py."""
import numpy as np
import pytest
import torch

from oracle import fbank_np, models_torch, speaker_engine_py as ref
from wespeaker_b200 import synthetic as syn
from wespeaker_b200.speaker_engine import SpeakerEngine, chunk_frame_index


@pytest.mark.parametrize("T,n", [(1, 198), (7, 198), (66, 198), (99, 198), (100, 198), (197, 198), (198, 198), (199, 198),
                                 (396, 198), (500, 198), (1000, 198), (5, 3), (3, 5), (0, 198)])
def test_chunk_index_matches_cpp_vector_logic(T, n):
    per_chunk_samples = 400 + (n - 1) * 160
    assert ref.num_chunk_frames(per_chunk_samples) == n
    want = ref.extract_feature(list(range(T)), per_chunk_samples)
    got = chunk_frame_index(T, n)
    assert got.shape == (len(want), n)
    assert [list(r) for r in got] == want


@pytest.mark.parametrize("nsamp", [9000, 32000, 50000, 100000])
def test_engine_host_logic_with_oracle_frontend_and_model(nsamp):
    """Whole SpeakerEngine with the CPU oracles injected for fbank and the model: isolates the chunk / CMN / average logic."""
    name = "ECAPA_TDNN_c512"
    sd = syn.make_state_dict(name, seed=0)
    fwd = lambda feats: models_torch.forward(name, sd, torch.as_tensor(feats, dtype=torch.float32))
    wav = syn.make_wavs(1, nsamp, seed=7)[0]
    fb = lambda w: torch.from_numpy(fbank_np.fbank(np.asarray(w[0], dtype=np.float32), window_type="hamming"))
    eng = SpeakerEngine(model=None, embedding_size=192, samples_per_chunk=32000, fbank_fn=fb, embed_fn=fwd)
    got = eng.extract_embedding(wav)
    frames = list(fbank_np.fbank(wav.astype(np.float32), window_type="hamming"))
    want = ref.extract_embedding(frames, 32000, lambda c: fwd(c[None])[0].numpy())
    assert got.shape == (192,)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5
    assert SpeakerEngine.cosine_similarity(got, want) == pytest.approx(1.0, abs=1e-6)
    assert SpeakerEngine.cosine_similarity(got, -want) == pytest.approx(0.0, abs=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("nsamp", [9000, 50000, 163840])
def test_engine_gpu_matches_oracle(nsamp):
    from wespeaker_b200.models import from_synthetic
    name = "ECAPA_TDNN_c512"
    m = from_synthetic(name, 0, precision="fp32")
    wav = syn.make_wavs(1, nsamp, seed=7)[0]
    got = SpeakerEngine(m, samples_per_chunk=32000).extract_embedding(torch.from_numpy(wav))
    sd = syn.make_state_dict(name, seed=0)
    fwd = lambda feats: models_torch.forward(name, sd, torch.as_tensor(feats, dtype=torch.float32))
    frames = list(fbank_np.fbank(wav.astype(np.float32), window_type="hamming"))
    want = ref.extract_embedding(frames, 32000, lambda c: fwd(c[None])[0].numpy())
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-4
    # full mode == plain extraction of the whole recording
    full = SpeakerEngine(m, samples_per_chunk=-1).extract_embedding(torch.from_numpy(wav))
    ref_full = ref.extract_embedding(frames, -1, lambda c: fwd(c[None])[0].numpy())
    assert np.linalg.norm(full - ref_full) / np.linalg.norm(ref_full) < 1e-4
