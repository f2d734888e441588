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
from .audio_utils import _ragged_offsets, _stats_ptr, fbank_batch, get_mel_filters  # noqa: F401  (fbank_batch re-exported)
from .feature_transforms.global_cmvn import cmvn_denormalize_cuda
from .plans import get_stft_plan, require_cuda, upload_small


def logmel_batch(waveforms: List, sample_rate: int = 24000, win_length: int = 1200, hop_length: int = 300,
                 n_fft: int = 2048, win_fn: callable = torch.hann_window, n_mels: int = 80, f_min: float = 20.0,
                 f_max: float = 8000.0, eps: float = 1e-5, cmvn_mean=None, cmvn_std=None, device=None, stats=None):
    """log(clamp(mel @ |STFT|, eps)) for a list of 1-D waveforms in [-1, 1] -> list of [1 + n_i // hop, n_mels]
    float32 CUDA tensors; optional fused global CMVN; ``stats`` (float64 CUDA [2, n_mels]) accumulates the sum and
    sum of squares of the features inside the extraction kernel (see ``global_cmvn_from_sums``)."""
    first = waveforms[0]
    if n_fft < 64 or n_fft > 4096 or n_fft & (n_fft - 1):
        raise ValueError(f"n_fft = {n_fft}: the CUDA front-end takes a power of two in [64, 4096] "
                         "(2048 runs the register-resident kernels, the rest a generic shared-memory FFT)")
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
                                     float(eps), _lib.ptr(mean_d), _lib.ptr(std_d), _stats_ptr(stats, n_mels, dev),
                                     _lib.ptr(out), _lib.stream_ptr(dev))
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
    feat = feat.cpu().numpy()  # [T, n_mels] numpy like the reference (data_utils.py:68)
    if target_length is not None:
        feat = _trim_or_pad(feat, target_length)
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


def global_cmvn_from_sums(stats: torch.Tensor, n_frames: int):
    """mean / std from the fused accumulators of ``logmel_batch(..., stats=)`` / ``fbank_batch(..., stats=)``
    (float64 [2, n]: sum and sum of squares over ``n_frames`` frames): the formulas of get_global_cmvn
    (examples/speech_synthesis/data_utils.py:210-213) evaluated in float64, returned as float32."""
    s = stats.detach().cpu().numpy().astype(np.float64)
    mean = s[0] / n_frames
    var = s[1] / n_frames - mean ** 2
    return {"mean": mean.astype(np.float32), "std": np.sqrt(np.maximum(var, 1e-10)).astype(np.float32)}


def _global_cmvn_from_paths(paths, output_path: Optional[Path] = None, device=None, batch_bytes: int = 256 << 20):
    """get_global_cmvn over an explicit, ordered list of .npy files.  The reference's accumulators are float32 running
    sums in file order (``mean_x += frames.sum(axis=0)``), so the result depends on that order in its last bits: the
    per-file terms come from ``s2st_utterance_sums`` (bit-identical to numpy's axis-0 reductions) and are added here
    in the given order, in float32, exactly like the reference does."""
    dev = require_cuda(device)
    lib = _lib.load()
    mean_x, mean_x2, n_frames = None, None, 0
    pending, pending_bytes = [], 0

    def flush():
        nonlocal mean_x, mean_x2, pending, pending_bytes
        if not pending:
            return
        n_cols = pending[0].shape[1]
        fo = np.zeros(len(pending) + 1, np.int32)
        fo[1:] = np.cumsum([f.shape[0] for f in pending])
        staged = torch.empty(int(fo[-1]), n_cols, dtype=torch.float32, pin_memory=True)
        np.concatenate(pending, out=staged.numpy())
        x = staged.to(dev, non_blocking=True)
        fo_d = upload_small(fo, dev)
        sums = torch.empty(len(pending), 2, n_cols, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.s2st_utterance_sums(len(pending), _lib.ptr(fo_d), n_cols, _lib.ptr(x), _lib.ptr(sums), _lib.stream_ptr(dev))
        _lib.check(rc, "s2st_utterance_sums")
        for cur in sums.cpu().numpy():  # file order; float32 adds like `mean_x += cur_mean_x`
            mean_x = cur[0].copy() if mean_x is None else mean_x + cur[0]
            mean_x2 = cur[1].copy() if mean_x2 is None else mean_x2 + cur[1]
        pending, pending_bytes = [], 0

    for p in paths:
        with open(p, "rb") as f:
            frames = np.load(f).squeeze()
        if frames.ndim != 2:
            raise ValueError(f"{p}: expected a [T, n_feat] feature matrix after squeeze(), got shape {frames.shape}")
        frames = np.ascontiguousarray(frames, dtype=np.float32)
        if pending and frames.shape[1] != pending[0].shape[1]:
            raise ValueError(f"{p}: {frames.shape[1]} feature columns, the files before it have {pending[0].shape[1]}")
        n_frames += frames.shape[0]
        pending.append(frames)
        pending_bytes += frames.nbytes
        if pending_bytes >= batch_bytes:
            flush()
    flush()
    if mean_x is None:
        raise TypeError("unsupported operand type(s) for /=: 'NoneType' and 'int'")  # what the reference raises on an empty directory
    mean_x = mean_x / np.float32(n_frames)
    mean_x2 = mean_x2 / np.float32(n_frames)
    var_x = mean_x2 - mean_x ** 2
    std_x = np.sqrt(np.maximum(var_x, 1e-10))
    if output_path is not None:
        with open(output_path, "wb") as f:
            np.savez(f, mean=mean_x, std=std_x)
    else:
        return {"mean": mean_x, "std": std_x}


def get_global_cmvn(feature_root: Path, output_path: Optional[Path] = None):
    """Drop-in for examples/speech_synthesis/data_utils.py:190-220: mean / std over every ``*.npy`` feature file of
    ``feature_root`` (files visited in ``Path.glob`` order like the reference), float32 accumulation semantics of the
    reference, saved as ``np.savez(mean=, std=)`` to ``output_path`` or returned."""
    return _global_cmvn_from_paths(list(Path(feature_root).glob("*.npy")), output_path)


def gcmvn_denormalize(x: torch.Tensor, mean, std) -> torch.Tensor:
    """x [B, T, C] * std + mean on the GPU."""
    dev = require_cuda(x.device)
    xd = x.detach().to(dev, torch.float32)
    mean_d = torch.as_tensor(mean).to(dev, torch.float32).contiguous()
    std_d = torch.as_tensor(std).to(dev, torch.float32).contiguous()
    assert mean_d.shape[0] == std_d.shape[0] == x.shape[-1]
    return cmvn_denormalize_cuda(xd, mean_d, std_d).to(x.device, x.dtype)
