"""Feature front-end oracle (test infrastructure only -- see oracle/__init__.py).

* ``logmel_spectrogram``  examples/speech_synthesis/data_utils.py:46-76 via
                          TTSSpectrogram (audio_utils.py:259-271, no phase) and
                          TTSMelScale (audio_utils.py:274-285)
* ``kaldi_fbank``         audio_utils.py:136-149 -> torchaudio.compliance.kaldi.fbank
                          with every option but num_mel_bins / sample_frequency at
                          its default (dither 0, povey window, pre-emphasis 0.97,
                          DC removal, snip_edges, power spectrum, log, 20 Hz..Nyquist)
* ``global_cmvn``         feature_transforms/global_cmvn.py:26-29
* ``utterance_cmvn``      feature_transforms/utterance_cmvn.py:29-40
* ``specaugment``         feature_transforms/specaugment.py:79-131 (masking; no time warp)
* ``global_cmvn_stats``   examples/speech_synthesis/data_utils.py:190-220
* ``gcmvn_denormalize``   fairseq/speech_generator_for_s2st.py:21-29
"""
import numpy as np

from .griffin_lim import stft
from .mel import kaldi_mel_banks, slaney_mel_filters

FLT_EPS = np.float32(1.1920928955078125e-07)


def logmel_spectrogram(wave, sample_rate=24000, win_length=1200, hop_length=300, n_fft=2048,
                       n_mels=80, f_min=20.0, f_max=8000.0, eps=1e-5, mel=None):
    """wave [n] float32 in [-1, 1] -> [1 + n // hop, n_mels] float32."""
    mag, _ = stft(wave, n_fft, win_length, hop_length)  # [F, T]
    if mel is None:
        mel = slaney_mel_filters(sample_rate, n_fft, n_mels, f_min, f_max)
    m = (mel.astype(np.float64) @ mag.astype(np.float64)).astype(np.float32)
    return np.log(np.maximum(m, np.float32(eps))).astype(np.float32).T


def kaldi_frame_params(sample_rate):
    win = int(sample_rate * 25.0 * 0.001)
    shift = int(sample_rate * 10.0 * 0.001)
    padded = 1 if win == 0 else 2 ** (win - 1).bit_length()
    return win, shift, padded


def povey_window(win):
    n = np.arange(win, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / (win - 1))  # periodic=False
    return np.power(hann.astype(np.float32), np.float32(0.85)).astype(np.float32)


def kaldi_fbank(wave, sample_rate, n_bins=80):
    """wave [n] float32 (int16-scaled) -> [m, n_bins] float32, m = 1 + (n - win)//shift."""
    wave = np.asarray(wave, np.float32).reshape(-1)
    win, shift, padded = kaldi_frame_params(sample_rate)
    n = wave.shape[0]
    if n < win:
        return np.zeros((0, n_bins), np.float32)
    m = 1 + (n - win) // shift
    idx = np.arange(win)[None, :] + shift * np.arange(m)[:, None]
    fr = wave[idx].astype(np.float64)
    fr = fr - fr.mean(axis=1, keepdims=True)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
    fr = fr - 0.97 * prev
    fr = fr * povey_window(win).astype(np.float64)[None, :]
    spec = np.fft.rfft(fr, n=padded, axis=1)
    power = (spec.real ** 2 + spec.imag ** 2)
    banks = kaldi_mel_banks(n_bins, padded, float(sample_rate)).astype(np.float64)
    e = (power @ banks.T).astype(np.float32)
    return np.log(np.maximum(e, FLT_EPS)).astype(np.float32)


def global_cmvn(x, mean, std):
    return np.divide(np.subtract(x, mean), std)


def gcmvn_denormalize(x, mean, std):
    return (np.asarray(x, np.float32) * np.asarray(std, np.float32) + np.asarray(mean, np.float32)).astype(np.float32)


def global_cmvn_stats(feature_list):
    """Sum / sum-of-squares accumulation in the features' dtype, like the reference."""
    sx = None
    sx2 = None
    n = 0
    for f in feature_list:
        f = np.asarray(f)
        n += f.shape[0]
        sx = f.sum(axis=0) if sx is None else sx + f.sum(axis=0)
        sx2 = (f ** 2).sum(axis=0) if sx2 is None else sx2 + (f ** 2).sum(axis=0)
    mean = sx / n
    var = sx2 / n - mean ** 2
    return {"mean": mean, "std": np.sqrt(np.maximum(var, 1e-10))}


def get_global_cmvn(file_arrays):
    """examples/speech_synthesis/data_utils.py:190-220, statement for statement, over the already loaded (and squeezed)
    arrays in the order the reference's glob visited them: in-place float32 running sums, in-place divisions."""
    mean_x, mean_x2, n_frames = None, None, 0
    for frames in file_arrays:
        frames = np.asarray(frames).squeeze()
        n_frames += frames.shape[0]
        cur_mean_x = frames.sum(axis=0)
        if mean_x is None:
            mean_x = cur_mean_x
        else:
            mean_x += cur_mean_x
        cur_mean_x2 = (frames ** 2).sum(axis=0)
        if mean_x2 is None:
            mean_x2 = cur_mean_x2
        else:
            mean_x2 += cur_mean_x2
    mean_x /= n_frames
    mean_x2 /= n_frames
    var_x = mean_x2 - mean_x ** 2
    return {"mean": mean_x, "std": np.sqrt(np.maximum(var_x, 1e-10))}


def utterance_cmvn(x, norm_means=True, norm_vars=True):
    """feature_transforms/utterance_cmvn.py:29-40, statement for statement (numpy float32 arithmetic)."""
    mean = x.mean(axis=0)
    square_sums = (x ** 2).sum(axis=0)
    if norm_means:
        x = np.subtract(x, mean)
    if norm_vars:
        var = square_sums / x.shape[0] - mean ** 2
        std = np.sqrt(np.maximum(var, 1e-10))
        x = np.divide(x, std)
    return x


def resize_rows_linear(src, dst_h, ipp=True):
    """cv2.resize(src, dsize=(src.shape[1], dst_h), interpolation=cv2.INTER_LINEAR) for float32 input (third-party:
    OpenCV 4.x, what specaugment.py:104-107 calls).  The width is unchanged, so the horizontal pass is the identity and
    every output row blends two source rows.  Two arithmetics exist in the field and both are pinned by fixtures
    (tests/golden/specaug_warp.npz, made with cv2 4.13):
      ipp=False  OpenCV's own code (imgproc/resize.cpp, cv2.ipp.setUseIPP(False) or a build without IPP):
                 fy = (float)((dy + 0.5) * scale - 0.5) with scale = 1 / (dst_h / src_h) in double, sy = floor(fy), float
                 weights (1 - fy, fy), dst = S0 * b0 + S1 * b1 with both products rounded (baseline SSE, no FMA);
      ipp=True   the x86-64 opencv-python wheels (IPP on by default -- what the reference gets from pip): source
                 position in double, weight rounded to float, dst = fma(S1 - S0, w, S0).  IPP is closed source; this
                 form was found by search and is bit-exact on every fixture.
    Source rows are clamped to the image in both."""
    src = np.asarray(src, np.float32)
    src_h = src.shape[0]
    out = np.empty((dst_h, src.shape[1]), np.float32)
    scale = 1.0 / (float(dst_h) / float(src_h))
    for dy in range(dst_h):
        if ipp:
            pos = (dy + 0.5) * (float(src_h) / float(dst_h)) - 0.5
            sy = int(np.floor(pos))
            w = np.float32(pos - sy)
        else:
            fy = np.float32((dy + 0.5) * scale - 0.5)
            sy = int(np.floor(fy))
            w = np.float32(fy - np.float32(sy))
        s0, s1 = min(max(sy, 0), src_h - 1), min(max(sy + 1, 0), src_h - 1)
        if ipp:
            d = src[s1] - src[s0]
            out[dy] = (d.astype(np.float64) * np.float64(w) + src[s0].astype(np.float64)).astype(np.float32)
        else:
            out[dy] = src[s0] * (np.float32(1.0) - w) + src[s1] * w
    return out


def specaugment(spectrogram, freq_mask_n=0, freq_mask_f=0, time_mask_n=0, time_mask_t=0, time_mask_p=0.0,
                mask_value=0.0, time_warp_w=0, ipp=True):
    """feature_transforms/specaugment.py:79-131: same RNG call sequence on numpy's global generator; the time warp
    (:96-110) uses the restatement of OpenCV's linear resize above."""
    import math
    distorted = spectrogram.copy()
    num_frames, num_freqs = spectrogram.shape
    if mask_value is None:
        mask_value = spectrogram.mean()
    if num_frames == 0 or num_freqs < freq_mask_f:
        return spectrogram
    if time_warp_w > 0 and 2 * time_warp_w < num_frames:
        w0 = np.random.randint(time_warp_w, num_frames - time_warp_w)
        w = np.random.randint(-time_warp_w + 1, time_warp_w)
        distorted = np.concatenate((resize_rows_linear(distorted[:w0], w0 + w, ipp),
                                    resize_rows_linear(distorted[w0:], num_frames - w0 - w, ipp)), axis=0)
    for _ in range(freq_mask_n):
        f = np.random.randint(0, freq_mask_f)
        f0 = np.random.randint(0, num_freqs - f)
        if f != 0:
            distorted[:, f0:f0 + f] = mask_value
    max_time_mask_t = min(time_mask_t, math.floor(num_frames * time_mask_p))
    if max_time_mask_t < 1:
        return distorted
    for _ in range(time_mask_n):
        t = np.random.randint(0, max_time_mask_t)
        t0 = np.random.randint(0, num_frames - t)
        if t != 0:
            distorted[t0:t0 + t, :] = mask_value
    return distorted
