"""Host-side mirror of the reference's ``fairseq/data/audio/audio_utils.py`` for the hot path.

Same public names and argument meaning; the arithmetic runs in the sm_100a CUDA library:

=========================  =================================================  ======================
here                       reference                                          CUDA entry point
=========================  =================================================  ======================
``get_window``             audio_utils.py:218-223                             (host, init only)
``get_mel_filters``        audio_utils.py:234-242 (librosa.filters.mel)       (host, init only)
``TTSSpectrogram``         audio_utils.py:245-271                             ``s2st_stft``
``TTSMelScale``            audio_utils.py:274-285                             ``s2st_mel_project``
``_get_torchaudio_fbank``  audio_utils.py:136-149                             ``s2st_fbank``
``get_fbank``              audio_utils.py:152-168                             ``s2st_fbank``
=========================  =================================================  ======================
"""
import math
from pathlib import Path
from typing import BinaryIO, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .plans import get_fbank_plan, get_stft_plan, require_cuda, upload_small

SF_AUDIO_FILE_EXTENSIONS = {".wav", ".flac", ".ogg"}


def get_window(window_fn: callable, n_fft: int, win_length: int) -> torch.Tensor:
    """window_fn(win_length) centred in n_fft samples (left pad = padding // 2)."""
    padding = n_fft - win_length
    assert padding >= 0
    out = torch.zeros(n_fft, dtype=torch.float32)
    out[padding // 2: padding // 2 + win_length] = window_fn(win_length).float()
    return out


def get_fourier_basis(n_fft: int) -> torch.Tensor:
    """[2 * (n_fft//2 + 1), n_fft]: cos rows then -sin rows of the DFT (kept for API parity; the CUDA
    path never materialises it -- it runs FFTs)."""
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float64)[:, None]
    n = torch.arange(n_fft, dtype=torch.float64)[None, :]
    ang = 2.0 * math.pi * torch.remainder(k * n, n_fft) / n_fft
    return torch.cat([torch.cos(ang), -torch.sin(ang)], dim=0).float()


def _slaney_hz_to_mel(f: torch.Tensor) -> torch.Tensor:
    step, knee = 200.0 / 3.0, 1000.0
    lin = f / step
    log = knee / step + torch.log(torch.clamp(f, min=1e-30) / knee) * (27.0 / math.log(6.4))
    return torch.where(f >= knee, log, lin)


def _slaney_mel_to_hz(m: torch.Tensor) -> torch.Tensor:
    step, knee = 200.0 / 3.0, 1000.0
    lin = m * step
    log = knee * torch.exp((m - knee / step) * (math.log(6.4) / 27.0))
    return torch.where(m >= knee / step, log, lin)


def get_mel_filters(sample_rate: int, n_fft: int, n_mels: int, f_min: float, f_max: float) -> torch.Tensor:
    """What ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` returns (Slaney scale, Slaney area
    normalisation), computed in float64 and cast to float32; librosa itself is not needed."""
    f64 = torch.float64
    bins = torch.linspace(0.0, sample_rate / 2.0, n_fft // 2 + 1, dtype=f64)
    lo = _slaney_hz_to_mel(torch.tensor(float(f_min), dtype=f64))
    hi = _slaney_hz_to_mel(torch.tensor(float(f_max), dtype=f64))
    edges = _slaney_mel_to_hz(lo + (hi - lo) * torch.arange(n_mels + 2, dtype=f64) / (n_mels + 1))
    left, centre, right = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    rising = (bins[None, :] - left) / (centre - left)
    falling = (right - bins[None, :]) / (right - centre)
    tri = torch.clamp(torch.minimum(rising, falling), min=0.0)
    return (tri * (2.0 / (right - left))).float()


def _stats_ptr(stats, n_cols, dev):
    if stats is None:
        return None
    assert stats.is_cuda and stats.dtype == torch.float64 and tuple(stats.shape) == (2, n_cols) and stats.is_contiguous()
    assert stats.device == dev, "the statistics accumulator must live on the extraction device"
    return _lib.ptr(stats)


def _ragged_offsets(lengths, hop, device):
    """frame_offsets (int32) and wave_offsets (int64) device tensors for waveforms of ``lengths``."""
    frames = [1 + n // hop for n in lengths]
    fo = np.zeros(len(lengths) + 1, np.int32)
    fo[1:] = np.cumsum(frames)
    wo = np.zeros(len(lengths) + 1, np.int64)
    wo[1:] = np.cumsum(lengths)
    return upload_small(fo, device), upload_small(wo, device), frames


class TTSSpectrogram(torch.nn.Module):
    """|STFT| (and phase) with reflect padding of n_fft//2: same interface as the reference module.

    forward(waveform [B, L]) -> magnitude [B, n_fft//2+1, T] (and phase), T = 1 + L // hop.
    """

    def __init__(self, n_fft: int, win_length: int, hop_length: int, window_fn: callable = torch.hann_window,
                 return_phase: bool = False) -> None:
        super().__init__()
        self.n_fft, self.win_length, self.hop_length = n_fft, win_length, hop_length
        self.return_phase = return_phase
        self.register_buffer("window", window_fn(win_length).float())

    def _plan(self, device):
        return get_stft_plan(device, self.n_fft, self.win_length, self.hop_length, 1, self.window)

    def forward(self, waveform: torch.Tensor):
        assert waveform.dim() == 2, "expected [B, L]"
        dev = require_cuda(waveform.device)
        B, L = waveform.shape
        if L <= self.n_fft // 2:
            raise RuntimeError(f"Padding size should be less than the corresponding input dimension, but got: "
                               f"padding ({self.n_fft // 2}, {self.n_fft // 2}) at dimension 1 of input {list(waveform.shape)}")
        x = waveform.detach().to(dev, torch.float32).contiguous()
        plan = self._plan(dev)
        fo, wo, frames = _ragged_offsets([L] * B, self.hop_length, dev)
        T, F = frames[0], self.n_fft // 2 + 1
        mag = torch.empty(B * T, F, dtype=torch.float32, device=dev)
        phase = torch.empty_like(mag) if self.return_phase else None
        with torch.cuda.device(dev):
            rc = _lib.load().s2st_stft(plan.handle, B, B * T, _lib.ptr(wo), _lib.ptr(fo), _lib.ptr(x), _lib.ptr(mag),
                                       _lib.ptr(phase), _lib.stream_ptr(dev))
        _lib.check(rc, "s2st_stft")
        # frame-major -> the reference's [B, F, T] (layout plumbing only)
        mag = mag.view(B, T, F).transpose(1, 2).to(waveform.device, waveform.dtype)
        if self.return_phase:
            return mag, phase.view(B, T, F).transpose(1, 2).to(waveform.device, waveform.dtype)
        return mag


class TTSMelScale(torch.nn.Module):
    """basis [n_mels, n_stft] @ specgram [..., n_stft, T] -- same interface as the reference module."""

    def __init__(self, n_mels: int, sample_rate: int, f_min: float, f_max: float, n_stft: int) -> None:
        super().__init__()
        self.n_mels, self.n_stft = n_mels, n_stft
        self.register_buffer("basis", get_mel_filters(sample_rate, (n_stft - 1) * 2, n_mels, f_min, f_max))

    def forward(self, specgram: torch.Tensor) -> torch.Tensor:
        dev = require_cuda(specgram.device)
        shape = specgram.shape
        assert shape[-2] == self.n_stft
        n_fft = (self.n_stft - 1) * 2
        plan = get_stft_plan(dev, n_fft, n_fft, n_fft // 4, self.n_mels, torch.ones(n_fft), mel=self.basis)
        x = specgram.detach().to(dev, torch.float32).reshape(-1, self.n_stft, shape[-1]).transpose(1, 2).contiguous()
        rows = x.shape[0] * x.shape[1]
        out = torch.empty(rows, self.n_mels, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().s2st_mel_project(plan.handle, rows, _lib.ptr(x), _lib.ptr(out), _lib.stream_ptr(dev))
        _lib.check(rc, "s2st_mel_project")
        out = out.view(x.shape[0], x.shape[1], self.n_mels).transpose(1, 2)
        return out.reshape(shape[:-2] + (self.n_mels, shape[-1])).to(specgram.device, specgram.dtype)


def convert_waveform(waveform, sample_rate: int, normalize_volume: bool = False, to_mono: bool = False,
                     to_sample_rate: Optional[int] = None):
    """The reference's convert_waveform (audio_utils.py:20-62) runs sox effects (``gain -n``, ``rate``, ``channels 1``)
    through torchaudio.  Down-mixing is a channel mean and is done here; volume normalisation and resampling are sox's
    own arithmetic and are not reproduced: requesting them raises ``NotImplementedError`` instead of silently returning
    something else."""
    if normalize_volume:
        raise NotImplementedError("normalize_volume (sox 'gain -n') is not provided")
    if to_sample_rate is not None and to_sample_rate != sample_rate:
        raise NotImplementedError(f"resampling {sample_rate} -> {to_sample_rate} Hz (sox 'rate') is not provided")
    if to_mono and waveform.shape[0] > 1:
        waveform = waveform.mean(axis=0, keepdims=True)
    return waveform, sample_rate


def get_waveform(path_or_fp: Union[str, BinaryIO], normalization: bool = True, mono: bool = True,
                 frames: int = -1, start: int = 0, always_2d: bool = True, output_sample_rate: Optional[int] = None,
                 normalize_volume: bool = False) -> Tuple[np.ndarray, int]:
    """16-bit WAV/FLAC/OGG reader with the reference's full signature (audio_utils.py:65-109).  soundfile is used when
    it is installed; without it 16-bit PCM WAV files are parsed directly (``io_utils.read_wav16``)."""
    if isinstance(path_or_fp, str):
        ext = Path(path_or_fp).suffix
        if ext not in SF_AUDIO_FILE_EXTENSIONS:
            raise ValueError(f"Unsupported audio format: {ext}")
    try:
        import soundfile as sf
    except ImportError:
        sf = None
    if sf is not None:
        waveform, sample_rate = sf.read(path_or_fp, dtype="float32", always_2d=True, frames=frames, start=start)
        waveform = waveform.T  # T x C -> C x T
    else:
        from .io_utils import read_wav16
        raw = path_or_fp.read() if hasattr(path_or_fp, "read") else path_or_fp
        try:
            waveform, sample_rate = read_wav16(raw)
        except ValueError as e:
            raise ImportError(f"Please install soundfile: pip install soundfile ({e})")
        n = waveform.shape[1]
        begin = start if start >= 0 else max(n + start, 0)
        waveform = waveform[:, begin: n if frames < 0 else begin + frames]
    waveform, sample_rate = convert_waveform(waveform, sample_rate, normalize_volume=normalize_volume, to_mono=mono,
                                             to_sample_rate=output_sample_rate)
    if not normalization:
        waveform = waveform * (2 ** 15)  # denormalised to 16-bit signed integers
    if not always_2d:
        waveform = waveform.squeeze(axis=0)
    return waveform, sample_rate


def fbank_batch(waveforms, sample_rate: int, n_bins: int = 80, cmvn_mean=None, cmvn_std=None, device=None, stats=None):
    """Kaldi fbank for a list of 1-D int16-scaled float waveforms (tensors on any device, or numpy).

    Returns a list of [m_i, n_bins] float32 CUDA tensors (m_i = 1 + (n_i - win) // shift, 0 if too short);
    optional fused global CMVN.  This is the batched (data-parallel) form of ``_get_torchaudio_fbank``.
    ``stats``: optional float64 CUDA tensor [2, n_bins]; the kernel adds (sum, sum of squares) of the features (before
    any CMVN) to it -- the accumulators of ``get_global_cmvn`` without a second pass over the corpus.
    """
    dev = require_cuda(device if device is not None else (waveforms[0].device if isinstance(waveforms[0], torch.Tensor) else None))
    plan = get_fbank_plan(dev, sample_rate, n_bins)
    waves = [torch.as_tensor(w).reshape(-1) for w in waveforms]
    lengths = [int(w.numel()) for w in waves]
    frames = [0 if n < plan.win else 1 + (n - plan.win) // plan.shift for n in lengths]
    fo = np.zeros(len(waves) + 1, np.int32)
    fo[1:] = np.cumsum(frames)
    wo = np.zeros(len(waves) + 1, np.int64)
    wo[1:] = np.cumsum(lengths)
    total = int(fo[-1])
    out = torch.empty(total, n_bins, dtype=torch.float32, device=dev)
    if total > 0:
        flat = torch.cat([w.to(dev, torch.float32) for w in waves]).contiguous()
        fo_d, wo_d = upload_small(fo, dev), upload_small(wo, dev)
        mean_d = None if cmvn_mean is None else torch.as_tensor(cmvn_mean).to(dev, torch.float32).contiguous()
        std_d = None if cmvn_std is None else torch.as_tensor(cmvn_std).to(dev, torch.float32).contiguous()
        with torch.cuda.device(dev):
            rc = _lib.load().s2st_fbank(plan.handle, len(waves), total, _lib.ptr(wo_d), _lib.ptr(fo_d), _lib.ptr(flat),
                                        _lib.ptr(mean_d), _lib.ptr(std_d), _stats_ptr(stats, n_bins, dev), _lib.ptr(out),
                                        _lib.stream_ptr(dev))
        _lib.check(rc, "s2st_fbank")
    return [out[fo[i]: fo[i + 1]] for i in range(len(waves))]


def _get_torchaudio_fbank(waveform: np.ndarray, sample_rate, n_bins=80) -> Optional[np.ndarray]:
    """Same contract as the reference helper: [1, n] (or [n]) int16-scaled numpy in, [m, n_bins] numpy out."""
    w = np.asarray(waveform, dtype=np.float32)
    if w.ndim == 2:
        assert w.shape[0] == 1, "expected a mono waveform"
        w = w[0]
    plan = get_fbank_plan(None, int(sample_rate), n_bins)
    assert 2 <= plan.win <= w.shape[0], f"choose a window size {plan.win} that is [2, {w.shape[0]}]"
    return fbank_batch([w], int(sample_rate), n_bins)[0].cpu().numpy()


def get_fbank(path_or_fp: Union[str, BinaryIO], n_bins=80) -> np.ndarray:
    """Mel-filter bank features of an audio file (Kaldi-compliant, int16-scaled input)."""
    waveform, sample_rate = get_waveform(path_or_fp, normalization=False)
    return _get_torchaudio_fbank(waveform, sample_rate, n_bins)
