"""Griffin-Lim vocoder oracle (test infrastructure only -- see oracle/__init__.py).

Restates, with FFTs instead of the reference's dense DFT-by-convolution:

* window / padding            fairseq/data/audio/audio_utils.py:218-223
* STFT magnitude + phase      audio_utils.py:259-271   (TTSSpectrogram.forward)
* window-sum-square           fairseq/models/text_to_speech/vocoder.py:71-82
* inverse STFT                vocoder.py:84-100        (GriffinLim.inverse)
* Griffin-Lim loop            vocoder.py:102-110       (GriffinLim.forward)
* pseudo-inverse mel          vocoder.py:24-46         (PseudoInverseMelScale)
* vocoder forward             vocoder.py:136-144       (GriffinLimVocoder.forward)

Float64 is used inside a transform; every value the reference keeps in a
float32 tensor (spectra, phases, waveforms) is rounded to float32 here too.
"""
import numpy as np

from .mel import slaney_mel_filters

TINY = np.float32(1.1754944e-38)  # vocoder.py:69


def hann_periodic(win_length):
    """torch.hann_window(win_length) (periodic=True default), float32."""
    n = np.arange(win_length, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)).astype(np.float32)


def padded_window(n_fft, win_length, window=None):
    """audio_utils.py:218-223: window centred in n_fft with zero padding."""
    w = hann_periodic(win_length) if window is None else np.asarray(window, np.float32)
    pad = n_fft - win_length
    assert pad >= 0
    out = np.zeros(n_fft, np.float32)
    out[pad // 2: pad // 2 + win_length] = w
    return out


def window_sum_square(n_frames, hop_length, win_length, n_fft, window=None):
    """vocoder.py:71-82, float32 accumulation in frame order like the reference."""
    w_sq = padded_window(n_fft, win_length, window) ** 2
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, np.float32)
    for i in range(n_frames):
        o = i * hop_length
        x[o: o + n_fft] += w_sq[: max(0, min(n_fft, n - o))]
    return x


def stft(wave, n_fft, win_length, hop_length, window=None):
    """audio_utils.py:259-271.  wave [L] float32 -> (mag, phase) each [n_fft//2+1, T]."""
    wave = np.asarray(wave, np.float32)
    half = n_fft // 2
    if wave.shape[-1] <= half:
        # F.pad(mode='reflect') raises when padding >= input length
        raise RuntimeError("reflect padding %d must be less than input length %d" % (half, wave.shape[-1]))
    xp = np.pad(wave.astype(np.float64), (half, half), mode="reflect")
    n_frames = 1 + (xp.shape[0] - n_fft) // hop_length
    w = padded_window(n_fft, win_length, window).astype(np.float64)
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
    spec = np.fft.rfft(xp[idx] * w[None, :], axis=1)  # [T, F]
    re = spec.real.astype(np.float32)
    im = spec.imag.astype(np.float32)
    mag = np.sqrt(re.astype(np.float64) ** 2 + im.astype(np.float64) ** 2).astype(np.float32)
    phase = np.arctan2(im, re).astype(np.float32)
    return mag.T, phase.T


def istft(mag, phase, n_fft, win_length, hop_length, window=None):
    """vocoder.py:84-100.  mag, phase [F, T] float32 -> wave [(T-1)*hop] float32.

    The reference's inverse basis is pinverse(N/H * fourier_basis) * window, i.e.
    (H/N) * window * irfft(.), with the imaginary parts of the DC and Nyquist
    rows ignored (their basis rows are identically zero); the later ``*= N/H``
    cancels the H/N.
    """
    mag = np.asarray(mag, np.float32)
    phase = np.asarray(phase, np.float32)
    n_frames = mag.shape[1]
    re = (mag * np.cos(phase)).astype(np.float32).astype(np.float64)
    im = (mag * np.sin(phase)).astype(np.float32).astype(np.float64)
    spec = (re + 1j * im).T  # [T, F]
    frames = np.fft.irfft(spec, n=n_fft, axis=1)  # ignores imag of DC / Nyquist
    w = padded_window(n_fft, win_length, window).astype(np.float64)
    frames = frames * w[None, :]
    n = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(n, np.float64)
    for t in range(n_frames):
        y[t * hop_length: t * hop_length + n_fft] += frames[t]
    wss = window_sum_square(n_frames, hop_length, win_length, n_fft, window)
    nz = wss > TINY
    y[nz] /= wss[nz].astype(np.float64)
    y = y[n_fft // 2:]
    y = y[: -(n_fft // 2)]
    return y.astype(np.float32)


def griffin_lim(mag, init_phase, n_iter, n_fft, win_length, hop_length, window=None):
    """vocoder.py:102-110 with the initial phase supplied by the caller."""
    wave = istft(mag, init_phase, n_fft, win_length, hop_length, window)
    for _ in range(n_iter):
        _, ph = stft(wave, n_fft, win_length, hop_length, window)
        wave = istft(mag, ph, n_fft, win_length, hop_length, window)
    return wave


def random_phase(shape):
    """vocoder.py:103: consumes the GLOBAL numpy RNG exactly like the reference."""
    return np.angle(np.exp(2j * np.pi * np.random.rand(*shape))).astype(np.float32)


def pinv_mel_basis(sample_rate, n_fft, n_mels, f_min, f_max):
    """vocoder.py:28-32: pinverse of the float32 Slaney filterbank -> [F, n_mels]."""
    fb = slaney_mel_filters(sample_rate, n_fft, n_mels, f_min, f_max)
    return np.linalg.pinv(fb.astype(np.float64)).astype(np.float32)


def inverse_mel(logmel, basis):
    """vocoder.py:34-46,141: exp, basis @ mel, clamp(min=0).  [T, n_mels] -> [F, T]."""
    mel = np.exp(np.asarray(logmel, np.float32)).astype(np.float32).T  # [n_mels, T]
    spec = basis.astype(np.float64) @ mel.astype(np.float64)
    return np.maximum(spec, 0.0).astype(np.float32)


def vocoder_forward(logmel, init_phase, n_iter, sample_rate=24000, win_length=1200, hop_length=300,
                    n_fft=2048, n_mels=80, f_min=20.0, f_max=8000.0, basis=None):
    """GriffinLimVocoder.forward (vocoder.py:136-144) for one utterance [T, n_mels]."""
    if basis is None:
        basis = pinv_mel_basis(sample_rate, n_fft, n_mels, f_min, f_max)
    mag = inverse_mel(logmel, basis)
    return griffin_lim(mag, init_phase, n_iter, n_fft, win_length, hop_length)


def spectral_convergence(wave, mag, n_fft, win_length, hop_length):
    """|| |STFT(wave)| - mag ||_F / || mag ||_F  (the acceptance metric of BASELINE.json)."""
    m, _ = stft(wave, n_fft, win_length, hop_length)
    num = np.linalg.norm(m.astype(np.float64) - np.asarray(mag, np.float64))
    den = np.linalg.norm(np.asarray(mag, np.float64))
    return float(num / max(den, 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
