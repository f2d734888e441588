"""Mel-cepstral distortion with DTW alignment: drop-in for the validation metric of the S2ST tasks
(``examples/s2s_trans/tasks/s2s_translation.py:414-552``; duplicated in ``s2s_translation_mtl.py`` and
``fairseq/tasks/text_to_speech.py``).

Same function names, arguments and return values.  The pairwise distance, the DTW recurrence and the back trace run in
the CUDA library (``s2st_rms_dist``, ``s2st_dtw``: one CTA per pair instead of O(M+N) rounds of torch launches and one
host synchronisation per path step); the 13-dimensional MFCC front-end is torchaudio's transform, as in the reference
(third-party arithmetic there as well; its 50 ms / 1200-point FFT is outside this library's 2048-point kernels).
"""
from typing import List, Optional

import torch
import torch.nn.functional as F

from . import _lib
from .plans import require_cuda


def batch_dynamic_time_warping(distance: torch.Tensor, shapes: Optional[torch.Tensor] = None):
    """full batched DTW without any constraints

    distance:  (batchsize, max_M, max_N) matrix
    shapes: (batchsize,) vector specifying (M, N) for each entry
    returns cumdist (float32), backptr (int32: 0=left, 1=up-left, 2=up), pathmap (int32), all (batchsize, max_M, max_N)
    """
    dev = require_cuda(distance.device if distance.is_cuda else None)
    d = distance.to(dev, torch.float32).contiguous()
    bsz, m, n = d.shape
    cumdist = torch.empty_like(d)
    backptr = torch.empty(d.shape, dtype=torch.int32, device=dev)
    pathmap = torch.empty(d.shape, dtype=torch.int32, device=dev)
    sh = None if shapes is None else torch.as_tensor(shapes).to(dev, torch.int64).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.load().s2st_dtw(bsz, m, n, _lib.ptr(d), _lib.ptr(sh), _lib.ptr(cumdist), _lib.ptr(backptr),
                                  _lib.ptr(pathmap), _lib.stream_ptr(dev))
    _lib.check(rc, "s2st_dtw")
    out_dev = distance.device
    return cumdist.to(out_dev), backptr.to(out_dev), pathmap.to(out_dev)


def compute_rms_dist(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """(m, n) root-mean-square distance matrix from (m, d) and (n, d) matrices"""
    dev = require_cuda(x1.device if x1.is_cuda else None)
    a, b = x1.to(dev, torch.float32).contiguous(), x2.to(dev, torch.float32).contiguous()
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1]
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2st_rms_dist(a.shape[0], b.shape[0], a.shape[1], _lib.ptr(a), _lib.ptr(b), _lib.ptr(out),
                                       _lib.stream_ptr(dev))
    _lib.check(rc, "s2st_rms_dist")
    return out.to(x1.device)


def compute_l2_dist(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """(m, n) squared L2 distance matrix from (m, d) and (n, d) matrices"""
    return compute_rms_dist(x1, x2).pow(2) * x1.size(1)


def get_divisor(pathmap, normalize_type):
    if normalize_type is None:
        return 1
    elif normalize_type == "len1":
        return pathmap.size(0)
    elif normalize_type == "len2":
        return pathmap.size(1)
    elif normalize_type == "path":
        return pathmap.sum().item()
    else:
        raise ValueError(f"normalize_type {normalize_type} not supported")


def batch_compute_distortion(y1: List[torch.Tensor], y2: List[torch.Tensor], sr, feat_fn, dist_fn, normalize_type):
    d, s, x1, x2 = [], [], [], []
    for cur_y1, cur_y2 in zip(y1, y2):
        assert cur_y1.ndim == 1 and cur_y2.ndim == 1
        cur_x1, cur_x2 = feat_fn(cur_y1), feat_fn(cur_y2)
        x1.append(cur_x1)
        x2.append(cur_x2)
        d.append(dist_fn(cur_x1, cur_x2))
        s.append(d[-1].size())
    max_m, max_n = max(ss[0] for ss in s), max(ss[1] for ss in s)
    d = torch.stack([F.pad(dd, (0, max_n - dd.size(1), 0, max_m - dd.size(0))) for dd in d])
    s = torch.LongTensor(s).to(d.device)
    cumdists, backptrs, pathmaps = batch_dynamic_time_warping(d, s)
    rets = []
    for (m, n), cur_x1, cur_x2, dist, cumdist, backptr, pathmap in zip(s, x1, x2, d, cumdists, backptrs, pathmaps):
        cumdist, backptr, pathmap = cumdist[:m, :n], backptr[:m, :n], pathmap[:m, :n]
        distortion = cumdist[-1, -1] / get_divisor(pathmap, normalize_type)
        rets.append((distortion, (cur_x1, cur_x2, dist, cumdist, backptr, pathmap)))
    return rets


def batch_mel_cepstral_distortion(y1, y2, sr, normalize_type="path", mfcc_fn=None):
    """
    https://arxiv.org/pdf/2011.03568.pdf

    The root mean squared error computed on 13-dimensional MFCC using DTW for
    alignment. MFCC features are computed from an 80-channel log-mel
    spectrogram using a 50ms Hann window and hop of 12.5ms.

    y1: list of waveforms
    y2: list of waveforms
    sr: sampling rate
    """
    try:
        import torchaudio
    except ImportError:
        raise ImportError("Please install torchaudio: pip install torchaudio")
    if mfcc_fn is None or mfcc_fn.sample_rate != sr:
        melkwargs = {"n_fft": int(0.05 * sr), "win_length": int(0.05 * sr), "hop_length": int(0.0125 * sr), "f_min": 20,
                     "n_mels": 80, "window_fn": torch.hann_window}
        mfcc_fn = torchaudio.transforms.MFCC(sr, n_mfcc=13, log_mels=True, melkwargs=melkwargs).to(y1[0].device)
    return batch_compute_distortion(y1, y2, sr, lambda y: mfcc_fn(y).transpose(-1, -2), compute_rms_dist, normalize_type)
