"""Mel-cepstral distortion with DTW alignment: drop-in for the validation metric of the S2ST tasks
(``examples/s2s_trans/tasks/s2s_translation.py:414-552``; duplicated in ``s2s_translation_mtl.py`` and
``fairseq/tasks/text_to_speech.py``).

Same function names, arguments and return values.  The pairwise distance, the DTW recurrence and the back trace run in
the CUDA library (``s2st_rms_dist``, ``s2st_dtw``: one CTA per pair instead of O(M+N) rounds of torch launches and one
host synchronisation per path step); the 13-dimensional MFCC front-end is torchaudio's transform, as in the reference
(third-party arithmetic there as well; its 50 ms / 1200-point FFT is outside this library's 2048-point kernels).
"""
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .plans import require_cuda, upload_small


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


_DIVISORS = {
    None: lambda pathmap, path_len: 1,
    "len1": lambda pathmap, path_len: pathmap.size(0),
    "len2": lambda pathmap, path_len: pathmap.size(1),
    "path": lambda pathmap, path_len: path_len,
}


def get_divisor(pathmap, normalize_type, path_len=None):
    """What the accumulated distance is divided by (s2s_translation.py:478-488): 1, the first / second sequence length,
    or the number of cells on the warping path (``path_len``: supplied by the batched caller from one device-side
    reduction; counted here when called on its own, like the reference does)."""
    if normalize_type not in _DIVISORS:
        raise ValueError(f"normalize_type {normalize_type} not supported")
    if normalize_type == "path" and path_len is None:
        path_len = int(pathmap.sum())
    return _DIVISORS[normalize_type](pathmap, path_len)


def batch_compute_distortion(y1: List[torch.Tensor], y2: List[torch.Tensor], sr, feat_fn, dist_fn, normalize_type):
    """Distortion of every (y1[b], y2[b]) pair after DTW alignment (s2s_translation.py:491-520), device-first:

    * the features of all pairs are concatenated once; with the library's own ``compute_rms_dist`` as ``dist_fn`` the
      zero-padded [bsz, max_M, max_N] distance batch is written by ONE kernel (``s2st_rms_dist_batch``) instead of one
      distance launch, one ``F.pad`` and one ``stack`` slot per pair (any other ``dist_fn`` is applied per pair and
      padded on the device);
    * one DTW launch for the batch (``s2st_dtw``), one reduction for all path lengths, ONE host synchronisation for
      the whole batch (the reference synchronises per pair through ``.item()``).

    Returns the reference's structure: ``[(distortion, (x1, x2, dist, cumdist, backptr, pathmap)), ...]`` with the
    matrices sliced to the pair's (M, N)."""
    if normalize_type not in _DIVISORS:
        raise ValueError(f"normalize_type {normalize_type} not supported")
    assert len(y1) == len(y2) and len(y1) > 0
    assert all(a.ndim == 1 and b.ndim == 1 for a, b in zip(y1, y2))
    x1 = [feat_fn(a) for a in y1]
    x2 = [feat_fn(b) for b in y2]
    dev = require_cuda(x1[0].device if x1[0].is_cuda else None)
    ms, ns = [int(x.shape[0]) for x in x1], [int(x.shape[0]) for x in x2]
    bsz, max_m, max_n = len(x1), max(ms), max(ns)
    if dist_fn is compute_rms_dist:
        cat1 = torch.cat([x.to(dev, torch.float32) for x in x1]).contiguous()
        cat2 = torch.cat([x.to(dev, torch.float32) for x in x2]).contiguous()
        off1 = upload_small(np.concatenate([[0], np.cumsum(ms)]).astype(np.int32), dev)
        off2 = upload_small(np.concatenate([[0], np.cumsum(ns)]).astype(np.int32), dev)
        dist = torch.empty(bsz, max_m, max_n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().s2st_rms_dist_batch(bsz, max_m, max_n, cat1.shape[1], _lib.ptr(cat1), _lib.ptr(cat2),
                                                 _lib.ptr(off1), _lib.ptr(off2), _lib.ptr(dist), _lib.stream_ptr(dev))
        _lib.check(rc, "s2st_rms_dist_batch")
    else:
        dist = torch.zeros(bsz, max_m, max_n, dtype=torch.float32, device=dev)
        for b, (a, c) in enumerate(zip(x1, x2)):
            dist[b, : ms[b], : ns[b]] = dist_fn(a, c).to(dev, torch.float32)
    shapes = upload_small(np.stack([ms, ns], axis=1).astype(np.int64), dev)
    cumdists, backptrs, pathmaps = batch_dynamic_time_warping(dist, shapes)
    last = cumdists[torch.arange(bsz, device=dev), shapes[:, 0] - 1, shapes[:, 1] - 1]
    path_lens = pathmaps.sum(dim=(1, 2)).tolist() if normalize_type == "path" else [None] * bsz  # the one synchronisation
    rets = []
    for b in range(bsz):
        m, n = ms[b], ns[b]
        pathmap = pathmaps[b, :m, :n]
        distortion = last[b] / get_divisor(pathmap, normalize_type, path_lens[b])
        rets.append((distortion, (x1[b], x2[b], dist[b], cumdists[b, :m, :n], backptrs[b, :m, :n], pathmap)))
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
