"""CPU oracle for the waveform-synthesis / feature front-end hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/`` (including the measurement scripts under ``tests/measure/`` that
time the reference's formulations), ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs may import it, and only as the checker (or
as the timed baseline), never as a fallback for the CUDA path.

It is a numpy restatement (FFT based, float64 inside a transform, float32 at
every point where the reference stores a float32 tensor) of the reference's
algorithms:

* ``oracle.mel``          -- librosa Slaney filterbank (third-party, unpinned in
                             the reference; call site audio_utils.py:234-242) and
                             Kaldi mel banks (torchaudio compliance/kaldi.py).
* ``oracle.griffin_lim``  -- vocoder.py:24-144 + audio_utils.py:218-271.
* ``oracle.frontend``     -- audio_utils.py:136-149,274-285,
                             examples/speech_synthesis/data_utils.py:46-76,190-220,
                             feature_transforms/global_cmvn.py:26-29, utterance_cmvn.py:29-40,
                             specaugment.py:79-131, speech_generator_for_s2st.py:21-29.
* ``oracle.conv_formulation`` -- the Griffin-Lim path once more, in the reference's own
                             dense-convolution formulation (torch; bit-identical to it).
* ``oracle.dtw``          -- examples/s2s_trans/tasks/s2s_translation.py:414-475 (MCD metric).

Parity pinning: the reference holds no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by ``tests/golden/make_golden.py``
(which imports the unmodified reference modules from /root/reference) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
the oracle against them on every run.
"""
