"""Mel filterbanks (oracle; test infrastructure only -- see oracle/__init__.py).

Two different third-party filterbanks sit on the reference's hot path:

* ``slaney_mel_filters``: what ``librosa.filters.mel(sr, n_fft, n_mels, fmin,
  fmax)`` (librosa < 0.10 positional signature, defaults htk=False,
  norm='slaney') returns.  The reference calls it at
  fairseq/data/audio/audio_utils.py:234-242 for both the vocoder's
  pseudo-inverse (vocoder.py:24-32) and the log-mel front-end
  (audio_utils.py:274-285).  librosa is not vendored and not pinned by the
  reference, so this restates its published recipe.
* ``kaldi_mel_banks``: torchaudio.compliance.kaldi.get_mel_banks without VTLN,
  used by ``fbank`` (reference call site audio_utils.py:141-147).
"""
import numpy as np

_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def _hz_to_slaney_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / _F_SP
    log = _MIN_LOG_MEL + np.log(np.maximum(f, 1e-30) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(f >= _MIN_LOG_HZ, log, lin)


def _slaney_mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * _F_SP
    log = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
    return np.where(m >= _MIN_LOG_MEL, log, lin)


def slaney_mel_filters(sample_rate, n_fft, n_mels, f_min, f_max):
    """[n_mels, n_fft//2+1] float32, Slaney scale, Slaney (area) normalisation."""
    n_freq = n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, sample_rate / 2.0, n_freq)
    edges_mel = np.linspace(_hz_to_slaney_mel(f_min), _hz_to_slaney_mel(f_max), n_mels + 2)
    edges_hz = _slaney_mel_to_hz(edges_mel)
    width = np.diff(edges_hz)
    ramps = edges_hz[:, None] - fft_freqs[None, :]
    fb = np.zeros((n_mels, n_freq), dtype=np.float64)
    for i in range(n_mels):
        rising = -ramps[i] / width[i]
        falling = ramps[i + 2] / width[i + 1]
        fb[i] = np.maximum(0.0, np.minimum(rising, falling))
    fb *= (2.0 / (edges_hz[2:] - edges_hz[:-2]))[:, None]
    return fb.astype(np.float32)


def kaldi_mel_banks(num_bins, padded_window_size, sample_freq, low_freq=20.0, high_freq=0.0):
    """[num_bins, padded//2 + 1] float32 (last column zero), as torchaudio's fbank
    builds it (compliance/kaldi.py get_mel_banks + the zero pad in fbank()).

    torchaudio evaluates this in float32 torch ops; the float32 evaluation order
    is reproduced here so the weights match to the last bit where possible.
    """
    f32 = np.float32
    num_fft_bins = padded_window_size // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded_window_size
    mel_low = 1127.0 * np.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * np.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = np.arange(num_bins, dtype=np.int64)[:, None]
    # python float (f64) scalar * int64 tensor -> float32 tensor in torch
    left = (f32(mel_low) + (b.astype(f32) * f32(delta))).astype(f32)
    center = (f32(mel_low) + ((b.astype(f32) + f32(1.0)) * f32(delta))).astype(f32)
    right = (f32(mel_low) + ((b.astype(f32) + f32(2.0)) * f32(delta))).astype(f32)
    freqs = (f32(fft_bin_width) * np.arange(num_fft_bins, dtype=f32)).astype(f32)
    mel = (f32(1127.0) * np.log((f32(1.0) + freqs / f32(700.0)).astype(f32)).astype(f32)).astype(f32)[None, :]
    up = ((mel - left) / (center - left)).astype(f32)
    down = ((right - mel) / (right - center)).astype(f32)
    bins = np.maximum(f32(0.0), np.minimum(up, down)).astype(f32)
    return np.concatenate([bins, np.zeros((num_bins, 1), f32)], axis=1)
