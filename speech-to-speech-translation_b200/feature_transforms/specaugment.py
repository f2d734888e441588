"""SpecAugment masking on the GPU: drop-in for ``fairseq/data/audio/feature_transforms/specaugment.py``
(``specaugment`` :12-131).

The mask rectangles are drawn on the host from numpy's global RNG with exactly the reference's call sequence
(``f``, ``f0`` per frequency mask :111-115, then ``t``, ``t0`` per time mask :123-127), so seeding numpy reproduces the
reference's masks; the fill -- and the "local mean" mask value when ``mask_value`` is None (:88-89) -- run in the
CUDA library.  ``apply_cuda`` masks a whole ragged, device-resident batch with one launch.  Time warping (:96-110,
``time_warp_W > 0``; the recipe leaves it at 0) draws ``w0`` and ``w`` like the reference and resizes the two parts of
the spectrogram with cv2.resize's float32 ``INTER_LINEAR`` arithmetic in ``s2st_time_warp`` (OpenCV itself is not needed).
"""
import math
import numbers
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib
from ..plans import require_cuda, upload_small
from . import AudioFeatureTransform, register_audio_feature_transform


def utterance_mean_cuda(x: torch.Tensor, frames: Sequence[int]) -> np.ndarray:
    """Mean of all elements of each utterance of a ragged batch x [sum T_i, n_feat] (float64 accumulation)."""
    fo = upload_small(np.concatenate([[0], np.cumsum(frames)]).astype(np.int32), x.device)
    sums = torch.zeros(len(frames), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().s2st_utterance_sum(len(frames), _lib.ptr(fo), x.shape[1], _lib.ptr(x), _lib.ptr(sums),
                                            _lib.stream_ptr(x.device))
    _lib.check(rc, "s2st_utterance_sum")
    counts = np.maximum(np.asarray(frames, np.float64) * x.shape[1], 1.0)
    return sums.cpu().numpy() / counts


# config key -> constructor argument, in the reference's order (specaugment.py:16-27)
_CONFIG_KEYS = (("time_warp_W", "time_warp_w", 0), ("freq_mask_N", "freq_mask_n", 0), ("freq_mask_F", "freq_mask_f", 0),
                ("time_mask_N", "time_mask_n", 0), ("time_mask_T", "time_mask_t", 0), ("time_mask_p", "time_mask_p", 0.0),
                ("mask_value", "mask_value", None))


@register_audio_feature_transform("specaugment")
class SpecAugmentTransform(AudioFeatureTransform):
    """SpecAugment (https://arxiv.org/abs/1904.08779): time warping, frequency and time masking.

    ``resize_arithmetic`` (instance or class attribute) selects which cv2.resize the time warp reproduces bit for bit:
    "ipp" = the x86-64 opencv-python wheels with IPP on (what ``pip install opencv-python`` gives the reference),
    "opencv" = OpenCV's own code (``cv2.ipp.setUseIPP(False)``, non-x86 builds).  They differ by ~1e-4 absolute."""

    resize_arithmetic = "ipp"

    @classmethod
    def from_config_dict(cls, config=None):
        cfg = config or {}
        return cls(**{arg: cfg.get(key, default) for key, arg, default in _CONFIG_KEYS})

    def __init__(self, time_warp_w: int = 0, freq_mask_n: int = 0, freq_mask_f: int = 0, time_mask_n: int = 0,
                 time_mask_t: int = 0, time_mask_p: float = 0.0, mask_value: Optional[float] = 0.0):
        # the reference's sanity checks (specaugment.py:40-53), same messages
        if not (mask_value is None or isinstance(mask_value, numbers.Number)):
            raise AssertionError(f"mask_value (type: {type(mask_value)}) must be None or a number")
        if freq_mask_n > 0 and not freq_mask_f > 0:
            raise AssertionError(f"freq_mask_F ({freq_mask_f}) must be larger than 0 when doing freq masking.")
        if time_mask_n > 0 and not time_mask_t > 0:
            raise AssertionError(f"time_mask_T ({time_mask_t}) must be larger than 0 when doing time masking.")
        for _key, arg, _default in _CONFIG_KEYS:
            setattr(self, arg, locals()[arg])

    def __repr__(self):
        shown = ", ".join(f"{arg}={getattr(self, arg)}" for _key, arg, _default in _CONFIG_KEYS[:-1])
        return f"{type(self).__name__}({shown})"

    def draw_warp(self, num_frames: int, num_freqs: int):
        """(w0, w) of the time warp for one spectrogram, consuming numpy's global RNG like specaugment.py:96-101, or
        (-1, 0) when the reference does not warp it."""
        if num_frames == 0 or num_freqs < self.freq_mask_f:
            return (-1, 0)
        if self.time_warp_w > 0 and 2 * self.time_warp_w < num_frames:
            w0 = np.random.randint(self.time_warp_w, num_frames - self.time_warp_w)
            w = np.random.randint(-self.time_warp_w + 1, self.time_warp_w)
            return (int(w0), int(w))
        return (-1, 0)

    def draw_masks(self, num_frames: int, num_freqs: int):
        """Rectangles (row0, row1, col0, col1) for one [num_frames, num_freqs] spectrogram, consuming numpy's global
        RNG exactly like specaugment.py:111-129.  None = the reference returns its input untouched."""
        if num_frames == 0 or num_freqs < self.freq_mask_f:
            return None
        rects = []
        for _i in range(self.freq_mask_n):
            f = np.random.randint(0, self.freq_mask_f)
            f0 = np.random.randint(0, num_freqs - f)
            if f != 0:
                rects.append((0, num_frames, f0, f0 + f))
        max_time_mask_t = min(self.time_mask_t, math.floor(num_frames * self.time_mask_p))
        if max_time_mask_t < 1:
            return rects
        for _i in range(self.time_mask_n):
            t = np.random.randint(0, max_time_mask_t)
            t0 = np.random.randint(0, num_frames - t)
            if t != 0:
                rects.append((t0, t0 + t, 0, num_freqs))
        return rects

    def apply_cuda(self, x: torch.Tensor, frames: Sequence[int]) -> torch.Tensor:
        """Mask a ragged batch x [sum T_i, n_feat] (float32, CUDA): utterances are visited in order, each drawing its
        masks like one reference call; returns a new tensor."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and sum(frames) == x.shape[0]
        x = x.contiguous()
        means = utterance_mean_cuda(x, frames) if self.mask_value is None else None  # of the un-warped input (:88-89)
        rects, values, warps, row = [], [], [], 0
        for i, T in enumerate(frames):
            warps.append(self.draw_warp(T, x.shape[1]))  # the reference draws the warp before the masks
            rs = self.draw_masks(T, x.shape[1])
            v = float(np.float32(means[i])) if self.mask_value is None else float(self.mask_value)
            for (r0, r1, c0, c1) in rs or []:
                rects.append((row + r0, row + r1, c0, c1))
                values.append(v)
            row += T
        if any(w0 >= 0 for w0, _ in warps):
            out = torch.empty_like(x)
            fo = upload_small(np.concatenate([[0], np.cumsum(frames)]).astype(np.int32), x.device)
            wd = upload_small(np.asarray(warps, np.int32), x.device)
            with torch.cuda.device(x.device):
                rc = _lib.load().s2st_time_warp(len(frames), x.shape[0], _lib.ptr(fo), x.shape[1], _lib.ptr(wd),
                                                {"ipp": 1, "opencv": 0}[self.resize_arithmetic], _lib.ptr(x),
                                                _lib.ptr(out), _lib.stream_ptr(x.device))
            _lib.check(rc, "s2st_time_warp")
        else:
            out = x.clone()
        if rects:
            rd = upload_small(np.asarray(rects, np.int32), x.device)
            vd = upload_small(np.asarray(values, np.float32), x.device)
            with torch.cuda.device(x.device):
                rc = _lib.load().s2st_fill_rects(len(rects), _lib.ptr(rd), _lib.ptr(vd), x.shape[1], _lib.ptr(out),
                                                 _lib.stream_ptr(x.device))
            _lib.check(rc, "s2st_fill_rects")
        return out

    def __call__(self, spectrogram):
        assert len(spectrogram.shape) == 2, "spectrogram must be a 2-D tensor."
        dev = require_cuda(None)
        if spectrogram.shape[0] == 0 or spectrogram.shape[1] < self.freq_mask_f:
            return spectrogram
        xd = torch.from_numpy(np.ascontiguousarray(spectrogram, np.float32)).to(dev)
        return self.apply_cuda(xd, [spectrogram.shape[0]]).cpu().numpy().astype(spectrogram.dtype, copy=False)
