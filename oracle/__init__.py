"""CPU oracle for the WeSpeaker embedding-extraction + PLDA hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the timed CPU
baseline), never as the path that is shipped or measured as the GPU number.

It is a restatement of the reference algorithms in plain torch-CPU fp32/fp64
functional ops (model forwards), numpy (fbank, CMN, PLDA) — each function cites the
reference file:line it follows.  It is *pinned* against the real reference code:
``tests/golden/make_golden.py`` imports the reference modules from /root/reference
(possible only in the build container), runs them on seeded inputs and commits the
outputs as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every oracle function against those fixtures.  fbank arithmetic is third-party
(``torchaudio.compliance.kaldi.fbank``, torchaudio 2.11.0 installed in this image;
the reference pins only ``torchaudio>=2.0.0`` in setup.py:35-36) and is pinned the
same way.
"""
