"""CPU oracle for the waveform-synthesis / feature front-end hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as a fallback for the CUDA path.

It is a numpy restatement (FFT based, float64 inside a transform, float32 at
every point where the reference stores a float32 tensor) of the reference's
algorithms:

* ``oracle.mel``          -- librosa Slaney filterbank (third-party, unpinned in
                             the reference; call site audio_utils.py:234-242) and
                             Kaldi mel banks (torchaudio compliance/kaldi.py).
* ``oracle.griffin_lim``  -- vocoder.py:24-144 + audio_utils.py:218-271.
* ``oracle.frontend``     -- audio_utils.py:136-149,274-285,
                             examples/speech_synthesis/data_utils.py:46-76,190-220,
                             feature_transforms/global_cmvn.py:26-29,
                             speech_generator_for_s2st.py:21-29.

Parity pinning: the reference holds no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by ``tests/golden/make_golden.py``
(which imports the unmodified reference modules from /root/reference) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
the oracle against them on every run.
"""
