"""Batched feature extraction (the data-parallel form of the reference's offline extractors).

* ``extract_fbank_features``       examples/speech_to_text/data_utils.py:73-98 (same signature)
* ``extract_logmel_spectrogram``   examples/speech_synthesis/data_utils.py:46-76 (same signature)
* ``logmel_batch`` / ``fbank_batch``  ragged batches, optional fused global CMVN
* ``global_cmvn_stats``            examples/speech_synthesis/data_utils.py:190-220
* ``gcmvn_denormalize``            fairseq/speech_generator_for_s2st.py:21-29
"""
from pathlib import Path
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .audio_utils import _ragged_offsets, fbank_batch, get_mel_filters  # noqa: F401  (fbank_batch re-exported)
from .feature_transforms.global_cmvn import cmvn_denormalize_cuda
from .plans import get_stft_plan, require_cuda, upload_small


def logmel_batch(waveforms: List, sample_rate: int = 24000, win_length: int = 1200, hop_length: int = 300,
                 n_fft: int = 2048, win_fn: callable = torch.hann_window, n_mels: int = 80, f_min: float = 20.0,
                 f_max: float = 8000.0, eps: float = 1e-5, cmvn_mean=None, cmvn_std=None, device=None):
    """log(clamp(mel @ |STFT|, eps)) for a list of 1-D waveforms in [-1, 1] -> list of [1 + n_i // hop, n_mels]
    float32 CUDA tensors; optional fused global CMVN."""
    first = waveforms[0]
    dev = require_cuda(device if device is not None else (first.device if isinstance(first, torch.Tensor) else None))
    mel = get_mel_filters(sample_rate, n_fft, n_mels, f_min, f_max)
    plan = get_stft_plan(dev, n_fft, win_length, hop_length, n_mels, win_fn(win_length), mel=mel)
    waves = [torch.as_tensor(w).reshape(-1) for w in waveforms]
    lengths = [int(w.numel()) for w in waves]
    for n in lengths:
        if n <= n_fft // 2:
            raise RuntimeError(f"Padding size should be less than the corresponding input dimension, but got: padding "
                               f"({n_fft // 2}, {n_fft // 2}) at dimension 2 of input [1, 1, {n}]")
    fo, wo, frames = _ragged_offsets(lengths, hop_length, dev)
    total = int(sum(frames))
    flat = torch.cat([w.to(dev, torch.float32) for w in waves]).contiguous()
    out = torch.empty(total, n_mels, dtype=torch.float32, device=dev)
    mean_d = None if cmvn_mean is None else torch.as_tensor(cmvn_mean).to(dev, torch.float32).contiguous()
    std_d = None if cmvn_std is None else torch.as_tensor(cmvn_std).to(dev, torch.float32).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.load().s2st_logmel(plan.handle, len(waves), total, _lib.ptr(wo), _lib.ptr(fo), _lib.ptr(flat),
                                     float(eps), _lib.ptr(mean_d), _lib.ptr(std_d), _lib.ptr(out), _lib.stream_ptr(dev))
    _lib.check(rc, "s2st_logmel")
    return list(torch.split(out, frames))


def _trim_or_pad(data: np.ndarray, target_length: int) -> np.ndarray:
    delta = data.shape[0] - target_length
    if delta >= 0:
        return data[:target_length]
    pad = np.zeros((-delta,) + data.shape[1:])
    return np.concatenate([data, pad], axis=0)


def extract_logmel_spectrogram(waveform: torch.Tensor, sample_rate: int, output_path: Optional[Path] = None,
                               win_length: int = 1024, hop_length: int = 256, n_fft: int = 1024,
                               win_fn: callable = torch.hann_window, n_mels: int = 80, f_min: float = 0.,
                               f_max: float = 8000, eps: float = 1e-5, overwrite: bool = False,
                               target_length: Optional[int] = None):
    if output_path is not None and output_path.is_file() and not overwrite:
        return
    assert waveform.dim() == 2 and waveform.shape[0] == 1
    feat = logmel_batch([waveform[0]], sample_rate, win_length, hop_length, n_fft, win_fn, n_mels, f_min, f_max, eps)[0]
    feat = feat.cpu()
    if target_length is not None:
        feat = _trim_or_pad(feat.numpy(), target_length)
    if output_path is not None:
        np.save(output_path.as_posix(), feat)
    else:
        return feat


def extract_fbank_features(waveform: torch.FloatTensor, sample_rate: int, output_path: Optional[Path] = None,
                           n_mel_bins: int = 80, overwrite: bool = False):
    if output_path is not None and output_path.is_file() and not overwrite:
        return
    w = waveform.mean(dim=0) if waveform.shape[0] > 1 else waveform[0]  # to mono
    w = w * (2 ** 15)  # Kaldi compliance: 16-bit signed integers
    features = fbank_batch([w], sample_rate, n_mel_bins)[0].cpu().numpy()
    if output_path is not None:
        np.save(output_path.as_posix(), features)
    return features


def global_cmvn_stats(features: List[torch.Tensor]):
    """mean / std over all frames of a list of [T_i, n_feat] CUDA tensors (sum and sum of squares on the GPU,
    accumulated in float64; std = sqrt(max(var, 1e-10)))."""
    dev = require_cuda(features[0].device)
    n_cols = features[0].shape[-1]
    sums = torch.zeros(2, n_cols, dtype=torch.float64, device=dev)
    n = 0
    lib = _lib.load()
    with torch.cuda.device(dev):
        for f in features:
            f = f.to(dev, torch.float32).contiguous()
            n += f.shape[0]
            rc = lib.s2st_cmvn_accumulate(f.shape[0], n_cols, _lib.ptr(f), _lib.ptr(sums), _lib.stream_ptr(dev))
            _lib.check(rc, "s2st_cmvn_accumulate")
    s = sums.cpu().numpy()
    mean = s[0] / n
    var = s[1] / n - mean ** 2
    return {"mean": mean.astype(np.float32), "std": np.sqrt(np.maximum(var, 1e-10)).astype(np.float32)}


def gcmvn_denormalize(x: torch.Tensor, mean, std) -> torch.Tensor:
    """x [B, T, C] * std + mean on the GPU."""
    dev = require_cuda(x.device)
    xd = x.detach().to(dev, torch.float32)
    mean_d = torch.as_tensor(mean).to(dev, torch.float32).contiguous()
    std_d = torch.as_tensor(std).to(dev, torch.float32).contiguous()
    assert mean_d.shape[0] == std_d.shape[0] == x.shape[-1]
    return cmvn_denormalize_cuda(xd, mean_d, std_d).to(x.device, x.dtype)
